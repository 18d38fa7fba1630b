// Stand-in for pcl::PointCloud<T> (a vector of points with PCL's forwarding members). TEST INFRASTRUCTURE ONLY.
#ifndef BALM_REF_STUB_PCL_POINT_CLOUD
#define BALM_REF_STUB_PCL_POINT_CLOUD
#include <memory>
#include <vector>
namespace pcl {
template <class T>
class PointCloud {
 public:
  typedef std::shared_ptr<PointCloud<T>> Ptr;
  typedef std::shared_ptr<const PointCloud<T>> ConstPtr;
  std::vector<T> points;
  unsigned width = 0, height = 1;
  void push_back(const T &p) { points.push_back(p); }
  size_t size() const { return points.size(); }
  void clear() { points.clear(); }
  void reserve(size_t n) { points.reserve(n); }
  void resize(size_t n) { points.resize(n); }
  bool empty() const { return points.empty(); }
  void swap(PointCloud &o) { points.swap(o.points); }
  T &operator[](size_t i) { return points[i]; }
  const T &operator[](size_t i) const { return points[i]; }
  typename std::vector<T>::iterator begin() { return points.begin(); }
  typename std::vector<T>::iterator end() { return points.end(); }
  typename std::vector<T>::const_iterator begin() const { return points.begin(); }
  typename std::vector<T>::const_iterator end() const { return points.end(); }
  PointCloud &operator+=(const PointCloud &o) { points.insert(points.end(), o.points.begin(), o.points.end()); return *this; }
};
}  // namespace pcl
#endif
