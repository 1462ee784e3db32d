/*
 * dcvc_oracle.cpp -- CPU oracle for stage 1 (semantic instance extraction).
 *
 * TEST INFRASTRUCTURE ONLY (see sgtd_oracle.h).  The clusterManager part is pinned
 * against the reference's own cluster_manager.hpp compiled into oracle/_ref
 * (tests/test_reference_build.py: label_info of every point and the clusters_ order);
 * gen_labels / gen_graphs are anchored on source lines only (get_json.cpp pulls in the
 * whole node: Python.h, ikd-Tree, fast_gicp, matplotlib).  Literal, per-point
 * restatement of
 *   gen_labels            R/src/get_json.cpp:41-229
 *   gen_graphs            R/src/get_json.cpp:231-343   (node part, :249-299)
 *   clusterManager        R/include/cluster_manager.hpp:139-421
 *     convert2polar :172-221, createHashTable/getPolarIndex :224-264,
 *     DCVC :272-355, searchKNN :365-385, labelAnalysis :394-421
 * including the O(N) relabel sweeps and the real std::unordered_map whose
 * iteration order decides the cluster (hence instance-id) order.
 * R = /root/reference/src/sgtd
 */
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <set>
#include <unordered_map>
#include <vector>

namespace {

struct DcvcParams {
  double startR, deltaR, deltaP, deltaA;
  int minSeg;
};

struct ClusterManager {
  DcvcParams params_;
  std::vector<std::array<float, 3>> selected_points_;
  double minPitch = 0.0, maxPitch = 0.0, minPolar = 5.0, maxPolar = 5.0; /* :482-485 */
  int width = 0, height = 0, polarNum = 0;
  std::vector<double> polarBounds;
  std::vector<std::array<double, 3>> polarCor;
  std::unordered_map<int, std::vector<int>> voxelMap;
  std::vector<std::vector<int>> clusters_; /* point indices (into selected_points_) per cluster */
  std::vector<int> labelInfo;

  /* convert2polar, :172-221 */
  void convert2polar() {
    auto azimuthCal = [](double x, double y) -> double {
      double angle = std::atan2(y, x);
      return angle > 0.0 ? angle * 180 / M_PI : (angle + 2 * M_PI) * 180 / M_PI;
    };
    size_t totalSize = selected_points_.size();
    /* std::vector<Eigen::Vector3d>::resize leaves new elements uninitialised in the
     * reference; freshly mapped memory is zero in practice (SURVEY 8a note v). */
    polarCor.assign(totalSize, {0.0, 0.0, 0.0});
    for (size_t i = 0; i < totalSize; ++i) {
      double cx = selected_points_[i][0], cy = selected_points_[i][1], cz = selected_points_[i][2];
      double r = std::sqrt(cx * cx + (cy * cy + cz * cz)); /* Eigen norm(): e0 + (e1 + e2), see sgtd_oracle.cpp norm3 */
      double pitch = std::asin(cz / r) * 180.0 / M_PI;
      double az = azimuthCal(cx, cy);
      if (r >= 120.0 || r <= 0.5) continue;
      minPitch = pitch < minPitch ? pitch : minPitch;
      maxPitch = pitch > maxPitch ? pitch : maxPitch;
      minPolar = r < minPolar ? r : minPolar;
      maxPolar = r > maxPolar ? r : maxPolar;
      polarCor[i] = {r, pitch, az};
    }
    polarNum = 0;
    polarBounds.clear();
    width = static_cast<int>(std::round(360.0 / params_.deltaA) + 1);
    height = static_cast<int>((maxPitch - minPitch) / params_.deltaP);
    double range = minPolar;
    int step = 1;
    while (range <= maxPolar) {
      range += (params_.startR - step * params_.deltaR);
      polarBounds.emplace_back(range);
      polarNum++, step++;
    }
  }
  int getPolarIndex(double radius) const { /* :259-264 */
    for (int r = 0; r < polarNum; ++r)
      if (radius < polarBounds[r]) return r;
    return polarNum - 1;
  }
  void indices(size_t item, int &polarIndex, int &pitchIndex, int &azimuthIndex, int &voxelIndex) const {
    const auto &cur = polarCor[item];
    polarIndex = getPolarIndex(cur[0]);
    pitchIndex = static_cast<int>(std::round((cur[1] - minPitch) / params_.deltaP));
    azimuthIndex = static_cast<int>(std::round(cur[2] / params_.deltaA));
    voxelIndex = (azimuthIndex * (polarNum + 1) + polarIndex) + pitchIndex * (polarNum + 1) * (width + 1);
  }
  void createHashTable() { /* :224-252 */
    size_t totalSize = polarCor.size();
    voxelMap.reserve(totalSize);
    for (size_t item = 0; item < totalSize; ++item) {
      int pi, pt, az, vi;
      indices(item, pi, pt, az, vi);
      auto iter = voxelMap.find(vi);
      if (iter != voxelMap.end()) iter->second.emplace_back((int)item);
      else voxelMap.insert(std::make_pair(vi, std::vector<int>{(int)item}));
    }
  }
  void searchKNN(int polar_index, int pitch_index, int azimuth_index, std::vector<int> &out) const { /* :365-385 */
    for (int z = pitch_index - 1; z <= pitch_index + 1; ++z) {
      if (z < 0 || z > height) continue;
      for (int y = polar_index - 1; y <= polar_index + 1; ++y) {
        if (y < 0 || y > polarNum) continue;
        for (int x = azimuth_index - 1; x <= azimuth_index + 1; ++x) {
          int ax = x;
          if (ax < 0) ax = width - 1;
          if (ax > 300) ax = 300;
          out.emplace_back((ax * (polarNum + 1) + y) + z * (polarNum + 1) * (width + 1));
        }
      }
    }
  }
  bool DCVC(std::vector<int> &label_info) { /* :272-355 */
    int labelCount = 0;
    size_t totalSize = polarCor.size();
    if (totalSize <= 0) return false;
    label_info.resize(totalSize, -1);
    for (size_t i = 0; i < totalSize; ++i) {
      if (label_info[i] != -1) continue;
      int polar_index, pitch_index, azimuth_index, voxel_index;
      indices(i, polar_index, pitch_index, azimuth_index, voxel_index);
      auto iter_find = voxelMap.find(voxel_index);
      std::vector<int> neighbors;
      if (iter_find != voxelMap.end()) {
        std::vector<int> KNN;
        searchKNN(polar_index, pitch_index, azimuth_index, KNN);
        for (auto &k : KNN) {
          iter_find = voxelMap.find(k);
          if (iter_find != voxelMap.end())
            for (auto &id : iter_find->second) neighbors.emplace_back(id);
        }
      }
      if (!neighbors.empty()) {
        for (auto &id : neighbors) {
          int currInfo = label_info[i];
          int neighInfo = label_info[id];
          if (currInfo != -1 && neighInfo != -1 && currInfo != neighInfo) {
            for (auto &seg : label_info)
              if (seg == currInfo) seg = neighInfo;
          } else if (neighInfo != -1) {
            label_info[i] = neighInfo;
          } else if (currInfo != -1) {
            label_info[id] = currInfo;
          } else {
            continue;
          }
        }
      }
      if (label_info[i] == -1) {
        labelCount++;
        label_info[i] = labelCount;
        for (auto &id : neighbors) label_info[id] = labelCount;
      }
    }
    return true;
  }
  void labelAnalysis(std::vector<int> &label_info) { /* :394-421 */
    std::unordered_map<int, std::vector<int>> label2segIndex;
    size_t totalSize = label_info.size();
    for (size_t i = 0; i < totalSize; ++i) label2segIndex[label_info[i]].emplace_back((int)i);
    for (auto &it : label2segIndex)
      if ((int)it.second.size() >= params_.minSeg) clusters_.push_back(it.second);
  }
  bool segmentPointCloud(const std::vector<std::array<float, 3>> &cloud) { /* :139-168 */
    selected_points_ = cloud;
    convert2polar();
    createHashTable();
    labelInfo.clear();
    if (!DCVC(labelInfo)) return false;
    labelAnalysis(labelInfo);
    return true;
  }
};

int node_map(int c) { /* R/src/get_json.cpp:10-12 ; -1 = key absent */
  switch (c) {
    case 10: return 3; case 11: return 4; case 12: return 5; case 13: return 6; case 14: return 7;
    case 15: return 8; case 16: return 9; case 17: return 10; case 18: return 11;
    case 0: case 1: case 2: case 3: case 4: case 5: case 6: case 7: case 8: case 19: return 0;
    default: return -1;
  }
}

} // namespace

extern "C" {

/* clusterManager::segmentPointCloud on one class cloud (for unit tests).
 * label_info[n]: DCVC label per point; cluster_of[n]: index of the point's cluster in
 * clusters_ order, -1 if its cluster is below minSeg.  Returns the number of clusters. */
int32_t orc_dcvc(const float *xyz, int64_t n, double startR, double deltaR, double deltaP, double deltaA,
                 int32_t minSeg, int32_t *label_info, int32_t *cluster_of, int32_t *grid /* width,height,polarNum */) {
  ClusterManager cm;
  cm.params_ = DcvcParams{startR, deltaR, deltaP, deltaA, minSeg};
  std::vector<std::array<float, 3>> cloud((size_t)n);
  for (int64_t i = 0; i < n; ++i) cloud[i] = {xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2]};
  for (int64_t i = 0; i < n; ++i) cluster_of[i] = -1;
  if (!cm.segmentPointCloud(cloud)) return 0;
  for (int64_t i = 0; i < n; ++i) label_info[i] = cm.labelInfo[i];
  for (size_t c = 0; c < cm.clusters_.size(); ++c)
    for (int idx : cm.clusters_[c]) cluster_of[idx] = (int32_t)c;
  if (grid) { grid[0] = cm.width; grid[1] = cm.height; grid[2] = cm.polarNum; }
  return (int32_t)cm.clusters_.size();
}

/* gen_labels + the node part of gen_graphs.
 * points: n x 4 float (x,y,z,intensity); labels: n x uint32 (lo16 semantic, hi16 instance).
 * point_instance[n]: instance id of each point, -1 if the point is in no instance.
 * node_xyz[cap*3], node_label[cap], node_inst[cap] (instance id of each node).
 * Returns 0, or -2 if cap_nodes is too small. */
static int32_t extract_instances_impl(const float *points, const uint32_t *labels, int64_t n, int32_t *point_instance,
                                      float *node_xyz, uint32_t *node_label, int32_t *node_inst, int32_t cap_nodes,
                                      int32_t *n_nodes, int32_t *n_instances, int variant) {
  std::vector<int> sem((size_t)n), ins((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    int lab = (int)labels[i];
    sem[i] = lab & 0xFFFF;
    ins[i] = lab >> 16;
    point_instance[i] = -1;
  }
  std::set<int> sem_set(sem.begin(), sem.end());
  int inst_id = 0;
  /* per instance: semantic label + original point indices in emission order */
  std::vector<std::pair<int, std::vector<int64_t>>> instances;
  for (int label_i : sem_set) {
    std::vector<int64_t> indice;
    for (int64_t i = 0; i < n; ++i)
      if (sem[i] == label_i) indice.push_back(i);
    std::set<int> inst_set;
    for (int64_t i : indice) inst_set.insert(ins[i]);
    if (label_i == 9 || label_i == 10) { /* :120-136 */
      instances.emplace_back(label_i, indice);
      inst_id += 1;
      continue;
    } else if (label_i == 0 || label_i == 1 || label_i == 2 || label_i == 3 || label_i == 6 || label_i == 7 ||
               label_i == 8 || label_i == 14 || (label_i == 19 && variant == 0)) {
      /* get_json.cpp:137; local_map.cpp:388 has the same list without 19 */
      continue;
    } else if (inst_set.size() > 1 || (inst_set.size() == 1 && *inst_set.begin() != 0)) { /* :138-159 */
      for (int label_j : inst_set) {
        std::vector<int64_t> pts;
        for (int64_t i : indice)
          if (ins[i] == label_j) pts.push_back(i);
        if (pts.size() <= 20) continue;
        instances.emplace_back(label_i, pts);
        inst_id += 1;
      }
    } else { /* :160-226 */
      int DCVC_min = 300;
      if (label_i == 17 || label_i == 18 || label_i == 15) DCVC_min = 5;
      /* local_map_creation (local_map.cpp:412-440): 400 for the large-surface classes */
      if (variant == 1 && (label_i == 10 || label_i == 11 || label_i == 12 || label_i == 14 || label_i == 16)) DCVC_min = 400;
      ClusterManager cm;
      cm.params_ = DcvcParams{0.35, 0.0004, 1.2, 1.2, DCVC_min};
      std::vector<std::array<float, 3>> cloud;
      cloud.reserve(indice.size());
      for (int64_t i : indice) cloud.push_back({points[i * 4], points[i * 4 + 1], points[i * 4 + 2]});
      cm.segmentPointCloud(cloud);
      for (const auto &cl : cm.clusters_) {
        std::vector<int64_t> pts;
        pts.reserve(cl.size());
        for (int idx : cl) pts.push_back(indice[(size_t)idx]);
        instances.emplace_back(label_i, pts);
        inst_id += 1;
      }
    }
  }
  *n_instances = inst_id;
  /* gen_graphs :249-299 : instances in ascending id */
  int32_t nn = 0;
  for (int id = 0; id < (int)instances.size(); ++id) {
    const int sem_label = instances[id].first;
    const auto &pts = instances[id].second;
    for (int64_t i : pts) point_instance[i] = id;
    const int mapped = node_map(sem_label);
    if (mapped < 0) continue; /* label 9: "else if (sem_label[0] == 9 || 10) continue" / unknown: skipped */
    float cx = 0.f, cy = 0.f, cz = 0.f; /* Eigen::Vector3f center_now += vec.head<3>() */
    for (int64_t i : pts) {
      cx += points[i * 4];
      cy += points[i * 4 + 1];
      cz += points[i * 4 + 2];
    }
    const float cnt = (float)pts.size(); /* center_now /= inst_cluster.size() */
    cx /= cnt; cy /= cnt; cz /= cnt;
    if (mapped >= 3 && mapped <= 12) {
      if (nn >= cap_nodes) return -2;
      node_xyz[nn * 3] = cx; node_xyz[nn * 3 + 1] = cy; node_xyz[nn * 3 + 2] = cz;
      node_label[nn] = (uint32_t)mapped;
      node_inst[nn] = id;
      ++nn;
    }
  }
  *n_nodes = nn;
  return 0;
}

int32_t orc_extract_instances(const float *points, const uint32_t *labels, int64_t n, int32_t *point_instance,
                              float *node_xyz, uint32_t *node_label, int32_t *node_inst, int32_t cap_nodes,
                              int32_t *n_nodes, int32_t *n_instances) {
  return extract_instances_impl(points, labels, n, point_instance, node_xyz, node_label, node_inst, cap_nodes, n_nodes,
                                n_instances, 0);
}
/* the same with the class tables of local_map_creation (R/src/local_map.cpp:384-440) */
int32_t orc_extract_instances_submap(const float *points, const uint32_t *labels, int64_t n, int32_t *point_instance,
                                     float *node_xyz, uint32_t *node_label, int32_t *node_inst, int32_t cap_nodes,
                                     int32_t *n_nodes, int32_t *n_instances) {
  return extract_instances_impl(points, labels, n, point_instance, node_xyz, node_label, node_inst, cap_nodes, n_nodes,
                                n_instances, 1);
}

/* The point-gathering part of local_map_creation (R/src/local_map.cpp:213-328), literally:
 * the scan's own points, then for every OTHER scan i of the submap whose translation lies within 15 m
 * (`(t1-t2).norm() > 15` skips) one transformed copy  T_j^-1 * T_i * BASE2OUSTER * p  -- of the CURRENT
 * scan's points again (the reference re-opens current_scan_path / current_label_path for every
 * neighbour, :272,:301), with the 4th component of each point (the intensity) acting as the homogeneous
 * coordinate except for the scan's last point, whose four components are set to one (:290 sets the last
 * COLUMN).  poses: nscans x 12 floats (row-major 3x4); b2o16: row-major 4x4.  Float arithmetic; the 4x4
 * inverse and products are Eigen's in the reference (order restated: cofactor inverse, k = 0..3 sums).
 * Returns the number of output points (n * copies), or -1 if cap is too small. */
static void mul44f(const float *a, const float *b, float *c) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float v = a[i * 4] * b[j];
      v += a[i * 4 + 1] * b[4 + j];
      v += a[i * 4 + 2] * b[8 + j];
      v += a[i * 4 + 3] * b[12 + j];
      c[i * 4 + j] = v;
    }
}
static void inv_rigid44f(const float *m, float *o) { /* general 4x4 with last row (0,0,0,1): cofactors of the 3x3 block */
  const float c00 = m[5] * m[10] - m[6] * m[9], c01 = m[6] * m[8] - m[4] * m[10], c02 = m[4] * m[9] - m[5] * m[8];
  const float det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  const float id = 1.0f / det;
  float r[9];
  r[0] = c00 * id; r[1] = (m[2] * m[9] - m[1] * m[10]) * id; r[2] = (m[1] * m[6] - m[2] * m[5]) * id;
  r[3] = c01 * id; r[4] = (m[0] * m[10] - m[2] * m[8]) * id; r[5] = (m[2] * m[4] - m[0] * m[6]) * id;
  r[6] = c02 * id; r[7] = (m[1] * m[8] - m[0] * m[9]) * id; r[8] = (m[0] * m[5] - m[1] * m[4]) * id;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) o[i * 4 + j] = r[i * 3 + j];
    o[i * 4 + 3] = -((r[i * 3] * m[3] + r[i * 3 + 1] * m[7]) + r[i * 3 + 2] * m[11]);
  }
  o[12] = o[13] = o[14] = 0.0f; o[15] = 1.0f;
}
int64_t orc_submap_aggregate(const float *points, const uint32_t *labels, int64_t n, const float *poses, int32_t nscans,
                             int32_t j, const float *b2o16, float radius, float *out_points, uint32_t *out_labels,
                             int64_t cap, int32_t *n_used) {
  auto pose44 = [&](int s, float *m) { for (int k = 0; k < 12; ++k) m[k] = poses[(size_t)s * 12 + k]; m[12] = m[13] = m[14] = 0.0f; m[15] = 1.0f; };
  float Tj[16], Tji[16];
  pose44(j, Tj);
  inv_rigid44f(Tj, Tji);
  int used = 1;
  int64_t o = 0;
  if (n > cap) return -1;
  for (int64_t p = 0; p < n; ++p) { for (int k = 0; k < 4; ++k) out_points[4 * o + k] = points[4 * p + k]; out_labels[o] = labels[p]; ++o; }
  for (int i = 0; i < nscans; ++i) {
    if (i == j) continue;
    const float dx = Tj[3] - poses[(size_t)i * 12 + 3], dy = Tj[7] - poses[(size_t)i * 12 + 7], dz = Tj[11] - poses[(size_t)i * 12 + 11];
    if (std::sqrt(dx * dx + (dy * dy + dz * dz)) > radius) continue; /* Eigen Vector3f::norm() */
    ++used;
    if (o + n > cap) return -1;
    float Ti[16], A[16], T[16];
    pose44(i, Ti);
    mul44f(Tji, Ti, A);
    mul44f(A, b2o16, T);
    for (int64_t p = 0; p < n; ++p) {
      float v[4] = {points[4 * p], points[4 * p + 1], points[4 * p + 2], points[4 * p + 3]};
      if (p == n - 1) v[0] = v[1] = v[2] = v[3] = 1.0f; /* points.col(points.cols() - 1).setOnes() */
      for (int r = 0; r < 4; ++r) {
        float acc = T[r * 4] * v[0];
        acc += T[r * 4 + 1] * v[1];
        acc += T[r * 4 + 2] * v[2];
        acc += T[r * 4 + 3] * v[3];
        out_points[4 * o + r] = acc;
      }
      out_labels[o] = labels[p];
      ++o;
    }
  }
  if (n_used) *n_used = used;
  return o;
}

} /* extern "C" */
