// oracle/shim/ros -- TEST INFRASTRUCTURE ONLY.  Stand-in for the ROS names in the
// reference's hot-path sources (ROS is not installed here).  NodeHandle::param is
// backed by a plain map so that read_parameters (R/src/STDesc.cpp:18-70) runs
// unmodified on the values of the reference's YAML.
#ifndef SGTD_SHIM_ROS
#define SGTD_SHIM_ROS
#include <bitset>
#include <chrono>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#define ROS_ERROR_STREAM(x) do { std::cerr << x << std::endl; } while (0)
#define ROS_ERROR(...) do { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } while (0)

namespace ros {
class NodeHandle {
 public:
  std::map<std::string, double> values;
  template <class T> bool param(const std::string &name, T &out, const T &dflt) const {
    auto it = values.find(name);
    if (it == values.end()) { out = dflt; return false; }
    out = (T)it->second;
    return true;
  }
};
class Publisher { public: template <class M> void publish(const M &) const {} };
}  // namespace ros

namespace geometry_msgs { struct Point { double x = 0, y = 0, z = 0; }; }
namespace visualization_msgs {
struct Marker {
  enum { LINE_LIST = 5, ADD = 0 };
  int type = 0, action = 0, id = 0;
  std::string ns;
  struct { double x = 0, y = 0, z = 0; } scale;
  struct { struct { double x = 0, y = 0, z = 0, w = 0; } orientation; } pose;
  struct { std::string frame_id; } header;
  struct { float r = 0, g = 0, b = 0, a = 0; } color;
  std::vector<geometry_msgs::Point> points;
};
struct MarkerArray { std::vector<Marker> markers; };
}  // namespace visualization_msgs
#endif
