// oracle/shim/pcl -- TEST INFRASTRUCTURE ONLY.
// Stand-in for the few PCL types the reference's hot-path sources name (PCL and FLANN
// are not installed here).  Written from scratch; see oracle/shim/Eigen/Core.
//
// pcl::KdTreeFLANN<PointT>::nearestKSearch restates what PCL 1.12 + FLANN 1.9 compute
// for this call (R/src/STDesc.cpp:183-192): exact k nearest neighbours under
// L2_Simple<float> (float accumulation ((dx*dx)+dy*dy)+dz*dz in x,y,z order), results
// sorted by ascending distance; k is clamped to the cloud size and the output vectors
// are resized to it (kdtree_flann.hpp).  FLANN breaks exact distance ties in tree
// traversal order; here ties go to the lower index (documented deviation).
#ifndef SGTD_SHIM_PCL
#define SGTD_SHIM_PCL
#include <stdint.h>

#include <algorithm>
#include <memory>
#include <vector>

#include "../Eigen/Core"

#define PCL_ADD_POINT4D float x, y, z, data_pad_;
#define PCL_ADD_INTENSITY float intensity
#define POINT_CLOUD_REGISTER_POINT_STRUCT(name, fields)

namespace boost { template <class T> using shared_ptr = std::shared_ptr<T>; }

namespace pcl {
struct PointXYZ { float x = 0, y = 0, z = 0; };
struct PointXYZI { float x = 0, y = 0, z = 0, intensity = 0; };
struct PointXYZL { float x = 0, y = 0, z = 0; uint32_t label = 0; };
struct PointXYZINormal { float x = 0, y = 0, z = 0, intensity = 0, normal_x = 0, normal_y = 0, normal_z = 0, curvature = 0; };
struct PointIndices { std::vector<int> indices; };

template <class PointT> class PointCloud {
 public:
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  std::vector<PointT> points;
  uint32_t width = 0, height = 0;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void push_back(const PointT &p) { points.push_back(p); }
  void clear() { points.clear(); }
  PointT &operator[](size_t i) { return points[i]; }
  const PointT &operator[](size_t i) const { return points[i]; }
  Ptr makeShared() const { return Ptr(new PointCloud<PointT>(*this)); }
};

template <class PointT> class KdTreeFLANN {
 public:
  typedef std::shared_ptr<KdTreeFLANN<PointT>> Ptr;
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr &cloud) { cloud_ = cloud; }
  int nearestKSearch(const PointT &p, int k, std::vector<int> &idx, std::vector<float> &dist) const {
    const int n = (int)cloud_->points.size();
    if (k > n) k = n;
    idx.resize(k); dist.resize(k);
    if (k == 0) return 0;
    std::vector<std::pair<float, int>> all(n);
    for (int i = 0; i < n; ++i) {
      const PointT &c = cloud_->points[i];
      const float dx = p.x - c.x, dy = p.y - c.y, dz = p.z - c.z;
      float d = dx * dx; d += dy * dy; d += dz * dz;
      all[i] = std::make_pair(d, i);
    }
    std::partial_sort(all.begin(), all.begin() + k, all.end());  // (distance, index) ascending
    for (int i = 0; i < k; ++i) { dist[i] = all[i].first; idx[i] = all[i].second; }
    return k;
  }
 private:
  typename PointCloud<PointT>::ConstPtr cloud_;
};

namespace search { template <class PointT> class KdTree { public: typedef std::shared_ptr<KdTree<PointT>> Ptr; }; }

// named by cluster_manager.hpp in never-instantiated templates / unused members only
template <class PointT> class VoxelGrid {
 public:
  template <class T> void setInputCloud(const T &) {}
  void setLeafSize(double, double, double) {}
  template <class T> void filter(T &) {}
};
template <class PointT> class EuclideanClusterExtraction {};
}  // namespace pcl
#endif
