/*
 * ref_wrap.cpp -- C interface of oracle/_ref/libsgtd_ref.so.
 *
 * TEST INFRASTRUCTURE ONLY.  libsgtd_ref.so is the REFERENCE's own hot-path code:
 *   /root/reference/src/sgtd/src/STDesc.cpp            (compiled unmodified, as its own TU)
 *   /root/reference/src/sgtd/include/desc/STDesc.h
 *   /root/reference/src/sgtd/include/cluster_manager.hpp (included below, unmodified)
 * built where those files lie (oracle/Makefile, target _ref) against the stand-in
 * headers of oracle/shim/ for Eigen / PCL / ROS / Ceres, none of which is installed in
 * this image.  What is the reference's: every line of BuildSingleScanSTD, AddSTDescs,
 * candidate_selector, candidate_verify, triangle_solver, SearchLoop, read_parameters,
 * Combinatorial_Binary_Encoding and clusterManager (convert2polar, createHashTable,
 * DCVC, searchKNN, labelAnalysis).  What is NOT (third-party arithmetic restated in
 * oracle/shim/): the exact-kNN of pcl::KdTreeFLANN, Eigen's JacobiSVD and Eigen's
 * small fixed-size expression evaluation order.
 *
 * This file only marshals plain arrays in and out of those functions; the layouts
 * (orc_desc, orc_cand) are the oracle's so tests can compare the two directly.
 * DB descriptors and query descriptors are tagged through the unused cov_mat_A_
 * field (global insertion index / query descriptor index) so that match and inlier
 * lists, which the reference returns as copies of STDesc pairs, can be reported as
 * index lists.
 */
#include <cstring>
#include <sstream>
/* every standard / shim header first, so that the access hack below only touches STDesc.h itself */
#include <ceres/ceres.h>
#include <pcl/common/io.h>
#include <ros/ros.h>
#include <Eigen/Core>
#include <fstream>
#include <mutex>
#include <unordered_map>
#include "omp.h"
#define private public /* candidate_selector / candidate_verify / triangle_solver are private members */
#include "desc/STDesc.h"
#undef private
#include "cluster_manager.hpp"

#include "sgtd_oracle.h"

int Combinatorial_Binary_Encoding(int a, int b, int c); /* R/src/STDesc.cpp:3-16 */

static std::string g_threads = "1";
extern "C" const char *ref_mp_proc_num(void) { return g_threads.c_str(); }

namespace {
struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };
/* candidate_selector prints "opm:<threads>" on every call (STDesc.cpp:322) */
struct QuietCout {
  NullBuf nb; std::streambuf *old;
  QuietCout() : old(std::cout.rdbuf(&nb)) {}
  ~QuietCout() { std::cout.rdbuf(old); }
};

struct Ref {
  STDescManager *m = nullptr;
  std::vector<STDesc> last;   /* result of the last BuildSingleScanSTD */
  int64_t n_db = 0;           /* descriptors added so far (next global insertion index) */
};

void to_orc(const STDesc &s, orc_desc *o) {
  memset(o, 0, sizeof(*o));
  for (int k = 0; k < 3; ++k) {
    o->side[k] = s.side_length_[k];
    o->vert[k] = (float)s.vertex_A_[k]; o->vert[3 + k] = (float)s.vertex_B_[k]; o->vert[6 + k] = (float)s.vertex_C_[k];
    o->lab[k] = (uint8_t)s.vertex_attached_[k];
  }
  o->frame = s.frame_id_;
  if (s.node_id.size() == 3) { o->anchor = (uint16_t)s.node_id[0]; o->m = (uint8_t)s.node_id[1]; o->n = (uint8_t)s.node_id[2]; }
}
STDesc from_orc(const orc_desc &o) {
  STDesc s;
  for (int k = 0; k < 3; ++k) {
    s.side_length_[k] = o.side[k];
    s.vertex_A_[k] = o.vert[k]; s.vertex_B_[k] = o.vert[3 + k]; s.vertex_C_[k] = o.vert[6 + k];
    s.vertex_attached_[k] = o.lab[k];
  }
  s.center_ = (s.vertex_A_ + s.vertex_B_ + s.vertex_C_) / 3; /* STDesc.cpp:296 */
  s.frame_id_ = o.frame;
  s.node_id = {o.anchor, o.m, o.n};
  return s;
}
}  // namespace

extern "C" {

void ref_set_threads(int n) { std::ostringstream os; os << (n < 1 ? 1 : n); g_threads = os.str(); }
int ref_max_frames(void) { return MAX_FRAME_N; }

/* ConfigSetting through the reference's own read_parameters (STDesc.cpp:18-70) */
void *ref_create(const orc_config *c) {
  ros::NodeHandle nh;
  nh.values["descriptor_near_num"] = c->descriptor_near_num;
  nh.values["candidate_num"] = c->candidate_num;
  nh.values["descriptor_min_len"] = c->descriptor_min_len;
  nh.values["descriptor_max_len"] = c->descriptor_max_len;
  nh.values["std_side_resolution"] = c->std_side_resolution;
  nh.values["rough_dis_threshold"] = c->rough_dis_threshold;
  nh.values["icp_threshold"] = c->icp_threshold;
  ConfigSetting cfg;
  QuietCout quiet;
  read_parameters(nh, cfg);
  Ref *r = new Ref;
  r->m = new STDescManager(cfg);
  return r;
}
void ref_destroy(void *h) { Ref *r = (Ref *)h; if (r) { delete r->m; delete r; } }
uint32_t ref_current_frame_id(const void *h) { return ((const Ref *)h)->m->current_frame_id_; }
int64_t ref_db_size(const void *h) { return ((const Ref *)h)->n_db; }
int ref_encode(int a, int b, int c) { return Combinatorial_Binary_Encoding(a, b, c); }

/* BuildSingleScanSTD (STDesc.cpp:174-315) on the node cloud Graph2CloudL would build. */
int64_t ref_build(void *h, const float *xyz, const uint32_t *label, int32_t K, orc_desc *out, int64_t cap) {
  Ref *r = (Ref *)h;
  pcl::PointCloud<pcl::PointXYZL>::Ptr pc(new pcl::PointCloud<pcl::PointXYZL>);
  for (int i = 0; i < K; ++i) {
    pcl::PointXYZL p; p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2]; p.label = label[i];
    pc->push_back(p);
  }
  r->m->BuildSingleScanSTD(pc, r->last);
  const int64_t n = (int64_t)r->last.size();
  for (int64_t i = 0; i < n && i < cap; ++i) to_orc(r->last[i], out + i);
  return n;
}

/* AddSTDescs (STDesc.cpp:149-172) of the descriptors ref_build just produced. */
void ref_add_last(void *h) {
  Ref *r = (Ref *)h;
  for (auto &s : r->last) s.cov_mat_A_(0, 0) = (double)(r->n_db++); /* tag: global insertion index */
  r->m->AddSTDescs(r->last);
}
/* AddSTDescs of descriptors given as plain records (one keyframe per call). */
void ref_add(void *h, const orc_desc *d, int64_t n) {
  Ref *r = (Ref *)h;
  std::vector<STDesc> v;
  for (int64_t i = 0; i < n; ++i) { v.push_back(from_orc(d[i])); v.back().cov_mat_A_(0, 0) = (double)(r->n_db++); }
  r->m->AddSTDescs(v);
}

/* SearchLoop (STDesc.cpp:84-147) on query descriptors q (NULL: the last ref_build result).
 * cands / m_q / m_g / inl as in orc_search; m_cell is not observable through the reference's
 * interface and best_hyp is reported as -1.  want_lists = 0 skips the extra candidate_selector
 * call that recovers the match lists (timing runs).  Returns the number of candidates, -1 on
 * empty input, -2 on capacity, -3 if the database has more keyframes than MAX_FRAME_N allows. */
int32_t ref_search(void *h, const orc_desc *q, int64_t nq, orc_cand *cands, int32_t cap_cand, int32_t *m_q,
                   uint32_t *m_g, int32_t *inl, int64_t cap_match, double best[2], int32_t want_lists) {
  Ref *r = (Ref *)h;
  if (r->m->current_frame_id_ >= (unsigned)MAX_FRAME_N) return -3; /* match_array[MAX_FRAME_N] would overflow */
  std::vector<STDesc> qs;
  if (q) for (int64_t i = 0; i < nq; ++i) qs.push_back(from_orc(q[i]));
  else qs = r->last;
  for (size_t i = 0; i < qs.size(); ++i) qs[i].cov_mat_A_(0, 0) = (double)i; /* tag: query descriptor index */
  QuietCout quiet;
  std::pair<int, double> loop_result(-1, 0);
  std::pair<Eigen::Vector3d, Eigen::Matrix3d> loop_transform;
  std::vector<std::pair<STDesc, STDesc>> loop_std_pair;
  std::vector<LOOP_RESULT> results;
  r->m->SearchLoop(qs, loop_result, loop_transform, loop_std_pair, results);
  best[0] = loop_result.first; best[1] = loop_result.second;
  if (qs.empty()) return -1;
  if ((int)results.size() > cap_cand) return -2;
  std::vector<STDMatchList> lists;
  if (want_lists) {
    r->m->candidate_selector(qs, lists);
    if (lists.size() != results.size()) return -4;
  }
  int64_t moff = 0, ioff = 0;
  for (size_t c = 0; c < results.size(); ++c) {
    const LOOP_RESULT &lr = results[c];
    orc_cand &o = cands[c];
    memset(&o, 0, sizeof(o));
    o.frame = lr.match_id; o.score = lr.match_fitness; o.best_hyp = -1;
    o.match_off = (int32_t)moff; o.inlier_off = (int32_t)ioff;
    o.ninlier = (int32_t)lr.loop_std_pair.size();
    for (int i = 0; i < 3; ++i) { o.t[i] = lr.loop_transform.first[i]; for (int j = 0; j < 3; ++j) o.R[i * 3 + j] = lr.loop_transform.second(i, j); }
    if (lr.match_fitness < 0) { o.R[0] = o.R[4] = o.R[8] = 1.0; o.R[1] = o.R[2] = o.R[3] = o.R[5] = o.R[6] = o.R[7] = 0; o.t[0] = o.t[1] = o.t[2] = 0; }
    if (want_lists) {
      const auto &ml = lists[c].match_list_;
      o.nmatch = o.votes = (int32_t)ml.size();
      if (moff + (int64_t)ml.size() > cap_match) return -2;
      for (size_t j = 0; j < ml.size(); ++j) {
        m_q[moff + j] = (int32_t)ml[j].first.cov_mat_A_(0, 0);
        m_g[moff + j] = (uint32_t)ml[j].second.cov_mat_A_(0, 0);
      }
      /* inliers are a subsequence of the match list (STDesc.cpp:523-538) */
      size_t p = 0;
      for (size_t j = 0; j < lr.loop_std_pair.size(); ++j) {
        const int qi = (int)lr.loop_std_pair[j].first.cov_mat_A_(0, 0);
        const uint32_t gi = (uint32_t)lr.loop_std_pair[j].second.cov_mat_A_(0, 0);
        while (p < ml.size() && !(m_q[moff + p] == qi && m_g[moff + p] == gi)) ++p;
        if (p == ml.size()) return -5;
        inl[ioff + j] = (int32_t)p++;
      }
      moff += (int64_t)ml.size();
    }
    ioff += o.ninlier;
  }
  return (int32_t)results.size();
}

/* triangle_solver (STDesc.cpp:549-571) on one pair. */
void ref_triangle_solver(const orc_desc *src, const orc_desc *ref, double R[9], double t[3]) {
  STDescManager m;
  std::pair<STDesc, STDesc> p(from_orc(*src), from_orc(*ref));
  Eigen::Vector3d tt; Eigen::Matrix3d rot;
  m.triangle_solver(p, tt, rot);
  for (int i = 0; i < 3; ++i) { t[i] = tt[i]; for (int j = 0; j < 3; ++j) R[i * 3 + j] = rot(i, j); }
}

/* clusterManager::segmentPointCloud (cluster_manager.hpp:139-169) on one class cloud, run step
 * by step (all members are public) so that DCVC's label_info is observable.  Same outputs as
 * orc_dcvc: label_info[n], cluster_of[n] (index in clusters_ order, -1 below minSeg), grid. */
int32_t ref_dcvc(const float *xyz, int64_t n, double startR, double deltaR, double deltaP, double deltaA,
                 int32_t minSeg, int32_t *label_info, int32_t *cluster_of, int32_t *grid) {
  clusterManager cm;
  clusterManager::ClusterParams params;
  clusterManager::DCVCParam seg;
  seg.startR = startR; seg.deltaR = deltaR; seg.deltaP = deltaP; seg.deltaA = deltaA; seg.minSeg = minSeg;
  cm.setParams(0, 0.5, 20, 2000, params, seg); /* as gen_labels does, get_json.cpp:187-197 */
  cm.reset(params);
  pcl::PointCloud<pcl::PointXYZ>::Ptr pc(new pcl::PointCloud<pcl::PointXYZ>);
  for (int64_t i = 0; i < n; ++i) { pcl::PointXYZ p; p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2]; pc->push_back(p); }
  for (int64_t i = 0; i < n; ++i) { cluster_of[i] = -1; label_info[i] = -1; }
  if (!cm.selectSemanticPoints(pc)) return 0;
  cm.convert2polar();
  cm.createHashTable();
  std::vector<int> labelInfo;
  if (!cm.DCVC(labelInfo)) return 0;
  if (grid) { grid[0] = cm.width; grid[1] = cm.height; grid[2] = cm.polarNum; }
  cm.labelAnalysis(labelInfo);
  for (int64_t i = 0; i < n; ++i) label_info[i] = labelInfo[i];
  /* clusters_ holds point copies in ascending point index; recover which label each one is */
  std::unordered_map<int, std::vector<int>> by_label;
  for (int64_t i = 0; i < n; ++i) by_label[labelInfo[i]].push_back((int)i);
  std::vector<char> used;
  std::vector<int> labs;
  for (auto &kv : by_label) if ((int)kv.second.size() >= minSeg) labs.push_back(kv.first);
  used.assign(labs.size(), 0);
  for (size_t c = 0; c < cm.clusters_.size(); ++c) {
    const auto &pts = cm.clusters_[c]->points;
    int found = -1;
    for (size_t l = 0; l < labs.size() && found < 0; ++l) {
      if (used[l]) continue;
      const std::vector<int> &idx = by_label[labs[l]];
      if (idx.size() != pts.size()) continue;
      bool same = true;
      for (size_t k = 0; k < idx.size() && same; ++k)
        same = pts[k].x == xyz[3 * idx[k]] && pts[k].y == xyz[3 * idx[k] + 1] && pts[k].z == xyz[3 * idx[k] + 2];
      if (same) found = (int)l;
    }
    if (found < 0) return -1;
    used[found] = 1;
    for (int i : by_label[labs[found]]) cluster_of[i] = (int32_t)c;
  }
  return (int32_t)cm.clusters_.size();
}

} /* extern "C" */
