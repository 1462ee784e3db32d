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
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

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

}  // extern "C"
