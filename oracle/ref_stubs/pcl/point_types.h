// Stand-in for the one PCL point type the reference uses (pcl::PointXYZINormal: x y z intensity normal_xyz curvature,
// with the data[4] alias of x,y,z that tools.hpp:213 reads). TEST INFRASTRUCTURE ONLY (see ../Eigen/Core).
#ifndef BALM_REF_STUB_PCL_POINT_TYPES
#define BALM_REF_STUB_PCL_POINT_TYPES
namespace pcl {
struct PointXYZINormal {
  union { float data[4]; struct { float x, y, z; }; };
  union { float data_n[4]; float normal[3]; struct { float normal_x, normal_y, normal_z; }; };
  float intensity = 0, curvature = 0;
  PointXYZINormal() { data[0] = data[1] = data[2] = 0; data[3] = 1; data_n[0] = data_n[1] = data_n[2] = data_n[3] = 0; }
};
}  // namespace pcl
#endif
