// sgtd/STDesc.h -- header-only C++ facade over the C ABI (include/sgtd_b200.h).
//
// Same names, members and call signatures as the reference's descriptor manager
// (R = /root/reference/src/sgtd):
//   ConfigSetting            R/include/desc/STDesc.h:38-72
//   STDesc                   R/include/desc/STDesc.h:75-97
//   LOOP_RESULT              R/include/desc/STDesc.h:99-104
//   STDescManager            R/include/desc/STDesc.h:342-440
//     BuildSingleScanSTD     R/src/STDesc.cpp:174-315
//     AddSTDescs             R/src/STDesc.cpp:149-172
//     SearchLoop             R/src/STDesc.cpp:84-147
//   read_parameters          R/src/STDesc.cpp:18-70
// so that R/src/semantic_graph_localization.cpp:415-417,457-458,592-602 compiles
// against it unchanged.  When Eigen / PCL / ROS headers are available they are
// used; otherwise small POD shims with the same spelling are provided (this
// image has none of them).  All computation happens in libsgtd_b200.so on the
// GPU; there is no CPU fallback.
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../sgtd_b200.h"

#if defined(__has_include)
#if __has_include(<Eigen/Core>)
#include <Eigen/Core>
#define SGTD_HAVE_EIGEN 1
#endif
#if __has_include(<pcl/point_cloud.h>) && __has_include(<pcl/point_types.h>)
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#define SGTD_HAVE_PCL 1
#endif
#if __has_include(<ros/ros.h>)
#include <ros/ros.h>
#define SGTD_HAVE_ROS 1
#endif
#endif

#ifndef SGTD_HAVE_EIGEN
// Minimal stand-ins with the accessors the node uses: operator[], operator(), <<-free.
namespace Eigen {
struct Vector3d {
  double v[3] = {0, 0, 0};
  double &operator[](int i) { return v[i]; }
  const double &operator[](int i) const { return v[i]; }
  double &operator()(int i) { return v[i]; }
  const double &operator()(int i) const { return v[i]; }
  double norm() const { return std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); }
};
struct Matrix3d {
  double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // row-major
  double &operator()(int r, int c) { return m[r * 3 + c]; }
  const double &operator()(int r, int c) const { return m[r * 3 + c]; }
  static Matrix3d Identity() { return Matrix3d(); }
};
}  // namespace Eigen
#endif

#ifndef SGTD_HAVE_PCL
namespace pcl {
struct PointXYZL {
  float x = 0, y = 0, z = 0;
  std::uint32_t label = 0;
};
template <typename PointT>
struct PointCloud {
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  std::vector<PointT> points;
  std::size_t size() const { return points.size(); }
  void push_back(const PointT &p) { points.push_back(p); }
};
}  // namespace pcl
#endif

typedef struct ConfigSetting {
  /* for point cloud pre-preocess*/
  int stop_skip_enable_ = 0;
  double ds_size_ = 0.5;
  int maximum_corner_num_ = 30;
  /* for key points*/
  double plane_merge_normal_thre_ = 0.1;
  double plane_merge_dis_thre_ = 0;
  double plane_detection_thre_ = 0.01;
  double voxel_size_ = 1.0;
  int voxel_init_num_ = 10;
  double proj_image_resolution_ = 0.5;
  double proj_dis_min_ = 0.2;
  double proj_dis_max_ = 5;
  double corner_thre_ = 10;
  /* for STD */
  int descriptor_near_num_ = 10;
  double descriptor_min_len_ = 1;
  double descriptor_max_len_ = 10;
  double non_max_suppression_radius_ = 3.0;
  double std_side_resolution_ = 0.2;
  /* for place recognition*/
  int skip_near_num_ = 50;
  int candidate_num_ = 50;
  int sub_frame_num_ = 10;
  double rough_dis_threshold_ = 0.03;
  double vertex_diff_threshold_ = 0.7;
  double icp_threshold_ = 0.5;
  double normal_threshold_ = 0.1;
  double dis_threshold_ = 0.3;
} ConfigSetting;

typedef struct STDesc {
  Eigen::Vector3d side_length_;
  Eigen::Vector3d angle_;
  Eigen::Vector3d center_;
  unsigned int frame_id_ = 0;
  Eigen::Vector3d vertex_A_, vertex_B_, vertex_C_;
  Eigen::Vector3d vertex_attached_;
  std::vector<int> node_id;
  Eigen::Matrix3d cov_mat_A_, cov_mat_B_, cov_mat_C_;
} STDesc;

struct LOOP_RESULT {
  int match_id;
  int match_fitness;
  std::pair<Eigen::Vector3d, Eigen::Matrix3d> loop_transform;
  std::vector<std::pair<STDesc, STDesc>> loop_std_pair;
};

namespace sgtd_detail {
inline sgtd_config to_c(const ConfigSetting &s) {
  sgtd_config c;
  std::memset(&c, 0, sizeof(c));
  c.stop_skip_enable = s.stop_skip_enable_; c.ds_size = s.ds_size_; c.maximum_corner_num = s.maximum_corner_num_;
  c.plane_merge_normal_thre = s.plane_merge_normal_thre_; c.plane_merge_dis_thre = s.plane_merge_dis_thre_;
  c.plane_detection_thre = s.plane_detection_thre_; c.voxel_size = s.voxel_size_; c.voxel_init_num = s.voxel_init_num_;
  c.proj_image_resolution = s.proj_image_resolution_; c.proj_dis_min = s.proj_dis_min_; c.proj_dis_max = s.proj_dis_max_;
  c.corner_thre = s.corner_thre_; c.descriptor_near_num = s.descriptor_near_num_; c.descriptor_min_len = s.descriptor_min_len_;
  c.descriptor_max_len = s.descriptor_max_len_; c.non_max_suppression_radius = s.non_max_suppression_radius_;
  c.std_side_resolution = s.std_side_resolution_; c.skip_near_num = s.skip_near_num_; c.candidate_num = s.candidate_num_;
  c.sub_frame_num = s.sub_frame_num_; c.rough_dis_threshold = s.rough_dis_threshold_;
  c.vertex_diff_threshold = s.vertex_diff_threshold_; c.icp_threshold = s.icp_threshold_;
  c.normal_threshold = s.normal_threshold_; c.dis_threshold = s.dis_threshold_;
  return c;
}
inline void from_c(const sgtd_config &c, ConfigSetting &s) {
  s.stop_skip_enable_ = c.stop_skip_enable; s.ds_size_ = c.ds_size; s.maximum_corner_num_ = c.maximum_corner_num;
  s.plane_merge_normal_thre_ = c.plane_merge_normal_thre; s.plane_merge_dis_thre_ = c.plane_merge_dis_thre;
  s.plane_detection_thre_ = c.plane_detection_thre; s.voxel_size_ = c.voxel_size; s.voxel_init_num_ = c.voxel_init_num;
  s.proj_image_resolution_ = c.proj_image_resolution; s.proj_dis_min_ = c.proj_dis_min; s.proj_dis_max_ = c.proj_dis_max;
  s.corner_thre_ = c.corner_thre; s.descriptor_near_num_ = c.descriptor_near_num; s.descriptor_min_len_ = c.descriptor_min_len;
  s.descriptor_max_len_ = c.descriptor_max_len; s.non_max_suppression_radius_ = c.non_max_suppression_radius;
  s.std_side_resolution_ = c.std_side_resolution; s.skip_near_num_ = c.skip_near_num; s.candidate_num_ = c.candidate_num;
  s.sub_frame_num_ = c.sub_frame_num; s.rough_dis_threshold_ = c.rough_dis_threshold;
  s.vertex_diff_threshold_ = c.vertex_diff_threshold; s.icp_threshold_ = c.icp_threshold;
  s.normal_threshold_ = c.normal_threshold; s.dis_threshold_ = c.dis_threshold;
}
inline STDesc to_std(const sgtd_desc &d) {
  STDesc s;
  for (int k = 0; k < 3; ++k) {
    s.side_length_[k] = d.side[k];
    s.vertex_A_[k] = d.vert[k]; s.vertex_B_[k] = d.vert[3 + k]; s.vertex_C_[k] = d.vert[6 + k];
    s.vertex_attached_[k] = d.lab[k];
    s.center_[k] = ((double)d.vert[k] + (double)d.vert[3 + k] + (double)d.vert[6 + k]) / 3;  // (A+B+C)/3
  }
  const double a = d.side[0], b = d.side[1], c = d.side[2];  // angle_ (STDesc.cpp:299-301); scale cancels
  s.angle_[0] = std::fabs((b * b + c * c - a * a) / (2 * b * c));
  s.angle_[1] = std::fabs((a * a + c * c - b * b) / (2 * a * c));
  s.angle_[2] = std::fabs((a * a + b * b - c * c) / (2 * a * b));
  s.frame_id_ = d.frame;
  s.node_id = {(int)d.anchor, (int)d.m, (int)d.n};
  return s;
}
inline sgtd_desc from_std(const STDesc &s) {
  sgtd_desc d;
  std::memset(&d, 0, sizeof(d));
  for (int k = 0; k < 3; ++k) {
    d.side[k] = s.side_length_[k];
    d.vert[k] = (float)s.vertex_A_[k]; d.vert[3 + k] = (float)s.vertex_B_[k]; d.vert[6 + k] = (float)s.vertex_C_[k];
    d.lab[k] = (std::uint8_t)(int)s.vertex_attached_[k];
  }
  d.frame = s.frame_id_;
  if (s.node_id.size() == 3) { d.anchor = (std::uint16_t)s.node_id[0]; d.m = (std::uint8_t)s.node_id[1]; d.n = (std::uint8_t)s.node_id[2]; }
  return d;
}
}  // namespace sgtd_detail

// read_parameters from the reference's YAML file (flat rosparam key names).
inline void read_parameters(const std::string &yaml_path, ConfigSetting &config_setting) {
  sgtd_config c;
  if (sgtd_config_from_yaml(yaml_path.c_str(), &c) != SGTD_OK) throw std::runtime_error("Error opening file: " + yaml_path);
  sgtd_detail::from_c(c, config_setting);
}
#ifdef SGTD_HAVE_ROS
inline void read_parameters(ros::NodeHandle &nh, ConfigSetting &c) {  // R/src/STDesc.cpp:18-56
  nh.param<double>("ds_size", c.ds_size_, 0.5);
  nh.param<int>("maximum_corner_num", c.maximum_corner_num_, 100);
  nh.param<double>("plane_merge_normal_thre", c.plane_merge_normal_thre_, 0.1);
  nh.param<double>("plane_detection_thre", c.plane_detection_thre_, 0.01);
  nh.param<double>("voxel_size", c.voxel_size_, 2.0);
  nh.param<int>("voxel_init_num", c.voxel_init_num_, 10);
  nh.param<double>("proj_image_resolution", c.proj_image_resolution_, 0.5);
  nh.param<double>("proj_dis_min", c.proj_dis_min_, 0);
  nh.param<double>("proj_dis_max", c.proj_dis_max_, 2);
  nh.param<double>("corner_thre", c.corner_thre_, 10);
  nh.param<int>("descriptor_near_num", c.descriptor_near_num_, 10);
  nh.param<double>("descriptor_min_len", c.descriptor_min_len_, 2);
  nh.param<double>("descriptor_max_len", c.descriptor_max_len_, 50);
  nh.param<double>("non_max_suppression_radius", c.non_max_suppression_radius_, 2.0);
  nh.param<double>("std_side_resolution", c.std_side_resolution_, 0.2);
  nh.param<int>("skip_near_num", c.skip_near_num_, 50);
  nh.param<int>("candidate_num", c.candidate_num_, 50);
  nh.param<int>("sub_frame_num", c.sub_frame_num_, 10);
  nh.param<double>("rough_dis_threshold", c.rough_dis_threshold_, 0.01);
  nh.param<double>("vertex_diff_threshold", c.vertex_diff_threshold_, 0.5);
  nh.param<double>("icp_threshold", c.icp_threshold_, 0.5);
  nh.param<double>("normal_threshold", c.normal_threshold_, 0.2);
  nh.param<double>("dis_threshold", c.dis_threshold_, 0.5);
}
#endif

class STDescManager {
 public:
  ConfigSetting config_setting_;
  int CS1 = 0;                         // probe-loop time of the last query, ms (vote kernel here)
  unsigned int current_frame_id_ = 0;

  explicit STDescManager(ConfigSetting &config_setting, int device = 0) : config_setting_(config_setting) {
    sgtd_config c = sgtd_detail::to_c(config_setting);
    if (sgtd_create(&c, device, &h_) != SGTD_OK) throw std::runtime_error(sgtd_last_error(nullptr));
    k_ = c.candidate_num;  // the handle keeps the values it was created with; config_setting_ stays editable like the reference's
  }
  ~STDescManager() { sgtd_destroy(h_); }
  STDescManager(const STDescManager &) = delete;
  STDescManager &operator=(const STDescManager &) = delete;
  sgtd_handle *handle() { return h_; }

  // generate STDescs from the instance-node cloud of one scan
  void BuildSingleScanSTD(const pcl::PointCloud<pcl::PointXYZL>::Ptr &instance_pc, std::vector<STDesc> &stds_vec) {
    stds_vec.clear();
    std::vector<sgtd_node> nodes(instance_pc->points.size());
    for (std::size_t i = 0; i < nodes.size(); ++i) {
      const auto &p = instance_pc->points[i];
      nodes[i] = sgtd_node{p.x, p.y, p.z, (std::uint32_t)p.label};
    }
    const std::int64_t off[2] = {0, (std::int64_t)nodes.size()};
    Batch b;
    // fewer nodes than descriptor_near_num: the reference reads stale kNN indices (UB); the library yields no descriptors
    check(sgtd_build_descriptors(h_, nodes.data(), off, 1, nullptr, &b.p));
    std::vector<sgtd_desc> d((std::size_t)sgtd_desc_batch_size(b.p));
    check(sgtd_desc_batch_download(h_, b.p, d.data(), nullptr));
    stds_vec.reserve(d.size());
    for (const auto &x : d) stds_vec.push_back(sgtd_detail::to_std(x));
  }

  // add descriptors of one keyframe to the database
  void AddSTDescs(const std::vector<STDesc> &stds_vec) {
    Batch b;
    upload(stds_vec, b);
    check(sgtd_add_descriptors(h_, b.p));
    current_frame_id_ = sgtd_current_frame_id(h_);
  }

  // search result <candidate_id, score>. -1 for no loop
  void SearchLoop(const std::vector<STDesc> &stds_vec, std::pair<int, double> &loop_result,
                  std::pair<Eigen::Vector3d, Eigen::Matrix3d> &loop_transform,
                  std::vector<std::pair<STDesc, STDesc>> &loop_std_pair, std::vector<LOOP_RESULT> &match_result_list) {
    if (stds_vec.empty()) {  // ROS_ERROR_STREAM("No STDescs!")
      loop_result = std::pair<int, double>(-1, 0);
      return;
    }
    Batch b;
    upload(stds_vec, b);
    Result res_guard;
    check(sgtd_search(h_, b.p, &res_guard.p));
    sgtd_search_result *r = res_guard.p;
    sgtd_loop_result lr;
    std::vector<sgtd_candidate> cands((std::size_t)k_);  // sized by the handle's candidate_num, not the editable field
    check(sgtd_result_download(h_, r, &lr, cands.data()));
    sgtd_timings tm;
    check(sgtd_result_stats(h_, r, nullptr, &tm));
    CS1 = (int)tm.vote_ms;
    int best = -1;
    for (int c = 0; c < lr.ncand; ++c) {
      const sgtd_candidate &cd = cands[c];
      LOOP_RESULT res;
      res.match_id = cd.frame;
      res.match_fitness = cd.score;
      for (int i = 0; i < 3; ++i) {
        res.loop_transform.first[i] = cd.t[i];
        for (int j = 0; j < 3; ++j) res.loop_transform.second(i, j) = cd.R[i * 3 + j];
      }
      if (cd.score > 0 && cd.match_off >= 0) {
        std::vector<std::int32_t> inl((std::size_t)cd.ninlier), mq((std::size_t)cd.nmatch);
        std::vector<std::uint32_t> mg((std::size_t)cd.nmatch), gsel((std::size_t)cd.ninlier);
        check(sgtd_result_inliers(h_, r, 0, c, inl.data(), cd.ninlier));
        check(sgtd_result_matches(h_, r, 0, c, mq.data(), nullptr, mg.data(), cd.nmatch));
        for (int i = 0; i < cd.ninlier; ++i) gsel[i] = mg[inl[i]];
        std::vector<sgtd_desc> dbd((std::size_t)cd.ninlier);
        check(sgtd_db_fetch(h_, gsel.data(), cd.ninlier, dbd.data()));
        res.loop_std_pair.reserve((std::size_t)cd.ninlier);
        for (int i = 0; i < cd.ninlier; ++i)
          res.loop_std_pair.emplace_back(stds_vec[(std::size_t)mq[inl[i]]], sgtd_detail::to_std(dbd[i]));
      }
      if (lr.frame >= 0 && best < 0 && cd.frame == lr.frame && (double)cd.score == lr.score) best = c;
      match_result_list.push_back(std::move(res));
    }
    if (lr.frame >= 0 && best >= 0) {
      const LOOP_RESULT &w = match_result_list[match_result_list.size() - (std::size_t)lr.ncand + (std::size_t)best];
      loop_result = std::pair<int, double>(lr.frame, lr.score);
      loop_transform = w.loop_transform;
      loop_std_pair = w.loop_std_pair;
    } else {
      loop_result = std::pair<int, double>(-1, 0);
    }
  }

 private:
  sgtd_handle *h_ = nullptr;
  int k_ = 0;  // candidate_num the handle was created with
  // device objects are released on every path, including the throwing ones
  struct Batch {
    sgtd_desc_batch *p = nullptr;
    Batch() = default;
    Batch(const Batch &) = delete;
    Batch &operator=(const Batch &) = delete;
    ~Batch() { if (p) sgtd_desc_batch_free(p); }
  };
  struct Result {
    sgtd_search_result *p = nullptr;
    Result() = default;
    Result(const Result &) = delete;
    Result &operator=(const Result &) = delete;
    ~Result() { if (p) sgtd_result_free(p); }
  };
  void check(int rc) {
    if (rc != SGTD_OK) throw std::runtime_error(sgtd_last_error(h_));
  }
  void upload(const std::vector<STDesc> &v, Batch &b) {
    std::vector<sgtd_desc> d(v.size());
    for (std::size_t i = 0; i < v.size(); ++i) d[i] = sgtd_detail::from_std(v[i]);
    const std::int64_t off[2] = {0, (std::int64_t)d.size()};
    check(sgtd_desc_batch_upload(h_, d.data(), off, 1, &b.p));
  }
};
