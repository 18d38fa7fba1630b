// Stand-in for the only ROS call inside the reference's BA header (ros::Time::now().toSec(), bavoxel.hpp:183,275).
#ifndef BALM_REF_STUB_ROS
#define BALM_REF_STUB_ROS
#include <chrono>
namespace ros {
struct Time {
  double s;
  static Time now() { return Time{std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count()}; }
  double toSec() const { return s; }
};
}  // namespace ros
#endif
