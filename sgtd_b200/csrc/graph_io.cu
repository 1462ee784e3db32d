// graph_io.cu -- host-side file formats around the hot path.
//
//   * the per-scan graph JSON that create_semantic_graph writes and the localization node
//     reads back: Graph::toJSON / fromJSON / readGraphFromFile
//     (R/include/Semantic_Graph.hpp:79-184).  Keys written: nodes, edges, weights, centers,
//     poses, volumes, densitys (the last four are always empty in the reference,
//     R/src/get_json.cpp:332); keys read: nodes, centers, poses (:147-151).
//   * KITTI .bin / .label scans as gen_labels reads them (R/src/get_json.cpp:47-84).
// float -> JSON double -> float round-trips exactly (the reference relies on the same).
// Plain C++ (no device code); lives in the library so a C++ host needs nothing else.
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include <algorithm>

#include "internal.cuh"

namespace {

// Minimal recursive-descent JSON reader: enough for objects / arrays / numbers / strings /
// literals; values of the three keys the reference reads are captured as flat number lists.
struct JsonReader {
  const char *p, *end;
  bool ok = true;
  void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
  bool lit(const char *s) { size_t n = strlen(s); if ((size_t)(end - p) >= n && !strncmp(p, s, n)) { p += n; return true; } return false; }
  std::string str() {
    std::string out;
    if (p >= end || *p != '"') { ok = false; return out; }
    ++p;
    while (p < end && *p != '"') { if (*p == '\\' && p + 1 < end) ++p; out.push_back(*p++); }
    if (p >= end) ok = false; else ++p;
    return out;
  }
  // parse any value; numbers met anywhere inside it are appended to `nums` (if non-null)
  void value(std::vector<double> *nums) {
    ws();
    if (p >= end) { ok = false; return; }
    if (*p == '{') {
      ++p; ws();
      if (p < end && *p == '}') { ++p; return; }
      while (ok) {
        ws(); str(); ws();
        if (p >= end || *p != ':') { ok = false; return; }
        ++p; value(nums); ws();
        if (p < end && *p == ',') { ++p; continue; }
        if (p < end && *p == '}') { ++p; return; }
        ok = false;
      }
    } else if (*p == '[') {
      ++p; ws();
      if (p < end && *p == ']') { ++p; return; }
      while (ok) {
        value(nums); ws();
        if (p < end && *p == ',') { ++p; continue; }
        if (p < end && *p == ']') { ++p; return; }
        ok = false;
      }
    } else if (*p == '"') {
      str();
    } else if (lit("true") || lit("false") || lit("null")) {
    } else {
      char *e = nullptr;
      double v = strtod(p, &e);
      if (e == p) { ok = false; return; }
      if (nums) nums->push_back(v);
      p = e;
    }
  }
};

}  // namespace

extern "C" {

// Graph(nodes, centers, poseRow).toJSON() -> file (R/src/get_json.cpp:332-341)
int sgtd_graph_write_json(const char *path, const sgtd_node *nodes, int32_t n_nodes, const float *poses12) {
  if (!path || (n_nodes > 0 && !nodes) || n_nodes < 0) return SGTD_E_INVALID;
  FILE *f = fopen(path, "w");
  if (!f) return SGTD_E_IO;
  // nlohmann::json orders object keys alphabetically
  fprintf(f, "{\"centers\":[");
  for (int i = 0; i < n_nodes; ++i)
    fprintf(f, "%s[%.17g,%.17g,%.17g]", i ? "," : "", (double)nodes[i].x, (double)nodes[i].y, (double)nodes[i].z);
  fprintf(f, "],\"densitys\":[],\"edges\":[],\"nodes\":[");
  for (int i = 0; i < n_nodes; ++i) fprintf(f, "%s%d", i ? "," : "", (int)nodes[i].label);
  fprintf(f, "],\"poses\":[");
  if (poses12) for (int i = 0; i < 12; ++i) fprintf(f, "%s%.17g", i ? "," : "", (double)poses12[i]);
  fprintf(f, "],\"volumes\":[],\"weights\":[]}");
  const bool bad = ferror(f);
  if (fclose(f) || bad) return SGTD_E_IO;
  return SGTD_OK;
}

// readGraphFromFile + fromJSON + Graph2CloudL: nodes/centers -> sgtd_node[], poses -> float[12]
// (R/include/Semantic_Graph.hpp:122-184, R/include/utility.hpp:646-659).
// *n_nodes receives the node count even when cap is too small (SGTD_E_CAPACITY).
int sgtd_graph_read_json(const char *path, sgtd_node *nodes, int32_t cap, int32_t *n_nodes, float *poses12,
                         int32_t *n_poses) {
  if (!path || !n_nodes) return SGTD_E_INVALID;
  std::ifstream in(path, std::ios::binary);
  if (!in.is_open()) return SGTD_E_IO;  // the reference throws std::runtime_error("Error opening file")
  std::stringstream ss;
  ss << in.rdbuf();
  const std::string text = ss.str();
  JsonReader r{text.data(), text.data() + text.size()};
  std::vector<double> v_nodes, v_centers, v_poses;
  r.ws();
  if (r.p >= r.end || *r.p != '{') return SGTD_E_INVALID;
  ++r.p; r.ws();
  bool have_nodes = false, have_centers = false;
  if (r.p < r.end && *r.p == '}') ++r.p;
  else
    while (r.ok) {
      r.ws();
      const std::string key = r.str();
      r.ws();
      if (r.p >= r.end || *r.p != ':') { r.ok = false; break; }
      ++r.p;
      if (key == "nodes") { r.value(&v_nodes); have_nodes = true; }
      else if (key == "centers") { r.value(&v_centers); have_centers = true; }
      else if (key == "poses") r.value(&v_poses);
      else r.value(nullptr);
      r.ws();
      if (r.p < r.end && *r.p == ',') { ++r.p; continue; }
      if (r.p < r.end && *r.p == '}') { ++r.p; break; }
      r.ok = false;
    }
  if (!r.ok || !have_nodes || !have_centers || v_centers.size() != 3 * v_nodes.size()) return SGTD_E_INVALID;
  *n_nodes = (int32_t)v_nodes.size();
  if (n_poses) *n_poses = (int32_t)v_poses.size();
  if (poses12) for (size_t i = 0; i < 12 && i < v_poses.size(); ++i) poses12[i] = (float)v_poses[i];
  if ((int32_t)v_nodes.size() > cap || (v_nodes.size() && !nodes)) return SGTD_E_CAPACITY;
  for (size_t i = 0; i < v_nodes.size(); ++i) {
    nodes[i].x = (float)v_centers[3 * i]; nodes[i].y = (float)v_centers[3 * i + 1]; nodes[i].z = (float)v_centers[3 * i + 2];
    nodes[i].label = (uint32_t)(int)v_nodes[i];
  }
  return SGTD_OK;
}

// KITTI scan: <scan>.bin = float32 x,y,z,intensity ; <scan>.label = uint32 (lo16 semantic, hi16 instance).
// Two-call pattern: points/labels may be NULL to query *n first.
int sgtd_scan_read_kitti(const char *bin_path, const char *label_path, float *points, uint32_t *labels, int64_t cap,
                         int64_t *n) {
  if (!bin_path || !n) return SGTD_E_INVALID;
  FILE *fb = fopen(bin_path, "rb");
  if (!fb) return SGTD_E_IO;
  fseek(fb, 0, SEEK_END);
  const int64_t npts = (int64_t)(ftell(fb) / 16);  // size / sizeof(float) / 4 (get_json.cpp:53-58)
  fseek(fb, 0, SEEK_SET);
  *n = npts;
  int rc = SGTD_OK;
  if (points) {
    if (npts > cap) rc = SGTD_E_CAPACITY;
    else if ((int64_t)fread(points, 16, (size_t)npts, fb) != npts) rc = SGTD_E_IO;
  }
  fclose(fb);
  if (rc || !label_path || !labels) return rc;
  FILE *fl = fopen(label_path, "rb");
  if (!fl) return SGTD_E_IO;
  fseek(fl, 0, SEEK_END);
  const int64_t nl = (int64_t)(ftell(fl) / 4);
  fseek(fl, 0, SEEK_SET);
  if (nl != npts) rc = SGTD_E_INVALID;  // assert(points.cols()==labels.size()) (get_json.cpp:77)
  else if (npts > cap) rc = SGTD_E_CAPACITY;
  else if ((int64_t)fread(labels, 4, (size_t)nl, fl) != nl) rc = SGTD_E_IO;
  fclose(fl);
  return rc;
}

// ---- evaluation helpers (host only) ----------------------------------------------------------
namespace {
// c = a * b for row-major 3x4 rigid/affine transforms with an implied last row (0, 0, 0, 1)
void mul34(const double *a, const double *b, double *c) {
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 4; ++j) {
      double v = (j == 3) ? a[i * 4 + 3] : 0.0;
      for (int k = 0; k < 3; ++k) v += a[i * 4 + k] * b[k * 4 + j];
      c[i * 4 + j] = v;
    }
  }
}
// inverse of a row-major 3x4 transform (general 3x3 block, not assumed orthonormal); false if singular
bool inv34(const double *a, double *o) {
  const double m00 = a[0], m01 = a[1], m02 = a[2], m10 = a[4], m11 = a[5], m12 = a[6], m20 = a[8], m21 = a[9], m22 = a[10];
  const double c00 = m11 * m22 - m12 * m21, c01 = m12 * m20 - m10 * m22, c02 = m10 * m21 - m11 * m20;
  const double det = m00 * c00 + m01 * c01 + m02 * c02;
  if (det == 0.0 || det != det) return false;
  const double id = 1.0 / det;
  double r[9];
  r[0] = c00 * id; r[1] = (m02 * m21 - m01 * m22) * id; r[2] = (m01 * m12 - m02 * m11) * id;
  r[3] = c01 * id; r[4] = (m00 * m22 - m02 * m20) * id; r[5] = (m02 * m10 - m00 * m12) * id;
  r[6] = c02 * id; r[7] = (m01 * m20 - m00 * m21) * id; r[8] = (m00 * m11 - m01 * m10) * id;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) o[i * 4 + j] = r[i * 3 + j];
    o[i * 4 + 3] = -(r[i * 3] * a[3] + r[i * 3 + 1] * a[7] + r[i * 3 + 2] * a[11]);
  }
  return true;
}
}  // namespace

int sgtd_pose_error(const double *gt12, const double *est12, double *t_err, double *r_err_deg) {
  if (!gt12 || !est12 || !t_err || !r_err_deg) return SGTD_E_INVALID;
  double inv[12], d[12];
  if (!inv34(est12, inv)) return SGTD_E_INVALID;
  mul34(inv, gt12, d);  // delta_T = lo.inverse() * gt (utility.hpp:114)
  *t_err = sqrt(d[3] * d[3] + d[7] * d[7] + d[11] * d[11]);
  const double c = fmin(fmax((d[0] + d[5] + d[10] - 1.0) / 2.0, -1.0), 1.0);
  *r_err_deg = fabs(acos(c)) / M_PI * 180.0;
  return SGTD_OK;
}

int sgtd_localization_check(const double *map_pose12, const double *R9, const double *t3, const double *refine12,
                            const double *gt12, const double *gt_extr12, double t_max, double r_max_deg,
                            double *est12, double *t_err, double *r_err_deg, int32_t *success) {
  if (!map_pose12 || !R9 || !t3 || !gt12 || !t_err || !r_err_deg || !success) return SGTD_E_INVALID;
  double loop[12], a[12], est[12], gt[12];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) loop[i * 4 + j] = R9[i * 3 + j];
    loop[i * 4 + 3] = t3[i];
  }
  mul34(map_pose12, loop, a);  // transform_j1 * new_trans (semantic_graph_localization.cpp:742)
  if (refine12) mul34(a, refine12, est);  // ... * transformation (the GICP refinement; identity when GICP is off)
  else memcpy(est, a, sizeof(est));
  if (gt_extr12) mul34(gt12, gt_extr12, gt);  // transform_test = transform_test * BASE2OUSTER (:741)
  else memcpy(gt, gt12, sizeof(gt));
  if (est12) memcpy(est12, est, sizeof(est));
  const int rc = sgtd_pose_error(gt, est, t_err, r_err_deg);
  if (rc != SGTD_OK) return rc;
  *success = (*t_err < t_max && *r_err_deg < r_max_deg) ? 1 : 0;  // :745
  return SGTD_OK;
}

// recall@k bookkeeping of the main loop (semantic_graph_localization.cpp:603-646): the candidates are
// ordered by match_fitness descending and walked until one lies within `radius` (10 m there) of the query's
// ground-truth pose; that position is the bin of STD_num[] that gets incremented.  The reference orders
// with std::sort (introsort, not stable), so the order among EQUAL fitness values is unspecified there;
// here ties keep candidate order (std::stable_sort).
int sgtd_recall_rank(const sgtd_candidate *cands, int32_t ncand, const double *map_poses12, int64_t n_map,
                     const double *gt12, double radius, int32_t *rank, int32_t *order) {
  if (!cands || ncand < 0 || !map_poses12 || !gt12 || !rank) return SGTD_E_INVALID;
  std::vector<int32_t> idx((size_t)ncand);
  for (int32_t i = 0; i < ncand; ++i) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) { return cands[a].score > cands[b].score; });
  *rank = -1;
  for (int32_t i = 0; i < ncand; ++i) {
    if (order) order[i] = idx[i];
    const int32_t f = cands[idx[i]].frame;
    if (*rank >= 0 || f < 0 || f >= n_map) continue;
    double t_e, r_e;
    // compute_adj_rpe(transform_t1 = query ground truth, transform1 = pose of keyframe match_id) (:637)
    if (sgtd_pose_error(gt12, map_poses12 + (size_t)f * 12, &t_e, &r_e) != SGTD_OK) continue;
    if (t_e < radius) *rank = i;
  }
  return SGTD_OK;
}

}  // extern "C"
