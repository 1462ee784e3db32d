// facade_node.cpp -- the reference node's call sequence against the facade
// (R/src/semantic_graph_localization.cpp:415-417 new manager, :455-458 map phase
//  Build+Add per keyframe, :590-603 query phase Build+SearchLoop), ROS-free.
// Input: a binary dump of node clouds (written by tests/test_facade_cpp.py).
// Output: one text line per query, compared with the oracle by the test.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "sgtd/STDesc.h"

static pcl::PointCloud<pcl::PointXYZL>::Ptr read_cloud(std::ifstream &in) {
  int32_t K = 0;
  in.read(reinterpret_cast<char *>(&K), 4);
  pcl::PointCloud<pcl::PointXYZL>::Ptr c(new pcl::PointCloud<pcl::PointXYZL>);
  for (int i = 0; i < K; ++i) {
    float xyz[3]; uint32_t lab;
    in.read(reinterpret_cast<char *>(xyz), 12);
    in.read(reinterpret_cast<char *>(&lab), 4);
    pcl::PointXYZL p; p.x = xyz[0]; p.y = xyz[1]; p.z = xyz[2]; p.label = lab;
    c->points.push_back(p);
  }
  return c;
}

#ifdef SGTD_HAVE_EIGEN
// The main loop's pose bookkeeping with the Eigen expressions the node uses on the facade's types
// (R/src/semantic_graph_localization.cpp:724-745, R/include/utility.hpp:110-123): comma initialiser, block<>
// assignment, cast<float>(), 4x4 products, inverse, topRightCorner, trace.
static void compute_adj_rpe(Eigen::Matrix4f &gt, Eigen::Matrix4f &lo, double &t_e, double &r_e) {
  Eigen::Matrix4f delta_T = lo.inverse() * gt;
  t_e = delta_T.topRightCorner(3, 1).norm();
  r_e = std::abs(std::acos(fmin(fmax((delta_T.block<3, 3>(0, 0).trace() - 1) / 2, -1.0), 1.0))) / M_PI * 180;
}
static void node_pose_check(const std::pair<Eigen::Vector3d, Eigen::Matrix3d> &loop_transform) {
  Eigen::Matrix4f new_trans = Eigen::Matrix4f::Identity();
  new_trans.block<3, 3>(0, 0) = loop_transform.second.cast<float>();
  new_trans.block<3, 1>(0, 3) = loop_transform.first.cast<float>();
  Eigen::Matrix3f Rot_test;
  Rot_test << 1, 0, 0, 0, 1, 0, 0, 0, 1;
  Eigen::Vector3f poses_test(0, 0, 0);
  Eigen::Matrix4f transform_test = Eigen::Matrix4f::Identity();
  transform_test.block<3, 3>(0, 0) = Rot_test;
  transform_test.block<3, 1>(0, 3) = poses_test;
  Eigen::Matrix4f transform_j1 = Eigen::Matrix4f::Identity(), transformation = Eigen::Matrix4f::Identity();
  Eigen::Matrix4f MAt_i1 = transform_j1 * new_trans * transformation;
  double T_error1, R_error1;
  compute_adj_rpe(transform_test, MAt_i1, T_error1, R_error1);
  // the same numbers through the library's helper (identity map pose and ground truth)
  const double I12[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  double R9[9], t3[3], te, re;
  int32_t ok;
  for (int i = 0; i < 3; ++i) { t3[i] = loop_transform.first[i]; for (int j = 0; j < 3; ++j) R9[i * 3 + j] = loop_transform.second(i, j); }
  sgtd_localization_check(I12, R9, t3, nullptr, I12, nullptr, 5.0, 10.0, nullptr, &te, &re, &ok);
  std::printf(" eigen_check %d", (std::fabs(te - T_error1) < 1e-3 * (1 + te) && std::fabs(re - R_error1) < 0.05) ? 1 : 0);
}
#endif

int main(int argc, char **argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: facade_node <yaml> <clouds.bin>\n"); return 2; }
  ConfigSetting config_setting;
  read_parameters(std::string(argv[1]), config_setting);
  STDescManager *std_manager = new STDescManager(config_setting);
  std::ifstream in(argv[2], std::ios::binary);
  int32_t nkf = 0, nq = 0;
  in.read(reinterpret_cast<char *>(&nkf), 4);
  in.read(reinterpret_cast<char *>(&nq), 4);
  for (int f = 0; f < nkf; ++f) {
    auto map_cloud = read_cloud(in);
    std::vector<STDesc> map_stds_vec;
    std_manager->BuildSingleScanSTD(map_cloud, map_stds_vec);
    std_manager->AddSTDescs(map_stds_vec);
  }
  std::printf("frames %u\n", std_manager->current_frame_id_);
  for (int q = 0; q < nq; ++q) {
    auto query_cloud = read_cloud(in);
    std::vector<STDesc> query_stds_vec;
    std_manager->BuildSingleScanSTD(query_cloud, query_stds_vec);
    std::pair<int, double> search_result(-1, 0);
    std::pair<Eigen::Vector3d, Eigen::Matrix3d> loop_transform;
    std::vector<std::pair<STDesc, STDesc>> loop_std_pair;
    std::vector<LOOP_RESULT> match_result_list;
    std_manager->SearchLoop(query_stds_vec, search_result, loop_transform, loop_std_pair, match_result_list);
    std::printf("query %d descs %zu best %d %.1f pairs %zu ncand %zu t %.9f %.9f %.9f R", q, query_stds_vec.size(),
                search_result.first, search_result.second, loop_std_pair.size(), match_result_list.size(),
                loop_transform.first[0], loop_transform.first[1], loop_transform.first[2]);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) std::printf(" %.12f", loop_transform.second(i, j));
#ifdef SGTD_HAVE_EIGEN
    if (search_result.first >= 0) node_pose_check(loop_transform);
#endif
    std::printf(" cands");
    for (auto &r : match_result_list) std::printf(" %d:%d:%zu", r.match_id, r.match_fitness, r.loop_std_pair.size());
    std::printf("\n");
  }
  delete std_manager;
  return 0;
}
