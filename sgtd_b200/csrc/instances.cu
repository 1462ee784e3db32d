// instances.cu -- stage 1: semantic instance extraction on sm_100a.
//
// Replaces gen_labels + the node part of gen_graphs (R/src/get_json.cpp:41-299) and
// clusterManager::segmentPointCloud (R/include/cluster_manager.hpp:139-421).
//
// The reference's DCVC labelling is sequential and order dependent (it is NOT
// connected components): points are visited in index order, only still-unlabelled
// points seed, a seed walks the points of its <= 27 neighbour curved voxels in a fixed
// order, neighbours met before the seed has a label stay unlabelled, a merge keeps
// the NEIGHBOUR's label value (cluster_manager.hpp:320-346).  Bit-exact membership
// therefore needs the sequential semantics.  What makes it tractable on a GPU:
//   * every point of a voxel has the same neighbour list, and a "visible" voxel
//     (pitch index <= height, so it is in its own neighbour list) is always in one of
//     three states -- NONE labelled, HEAD only (its lowest point seeded while the walk
//     had passed the voxel unlabelled), ALL same label -- so it produces at most two
//     seed events (its two lowest points);  points of "invisible" voxels (pitch index
//     = height+1, skipped by every walk including their own) seed individually;
//   * the O(N) relabel sweep is a union-find link  parent[cur] = neigh.
// So: a fully parallel pass (polar transform, curved-voxel open-addressing table with
// atomicCAS/atomicMin, event list by ordered compaction), a second one that writes every voxel's 27-neighbour
// row, and then the replay of the event list per (scan, class) task.  Events of different connected
// components of the voxel-neighbour graph commute, so the default replay (k_dcvc_replay_cc) finds the
// components with a 16-bit union-find and lets the 8 warps of the task's CTA replay them concurrently; new
// labels are named after the event that creates them and renamed to the reference's counter values by a
// prefix sum afterwards.  Oversized tasks use the sequential forms (k_dcvc_replay: one replaying warp, lanes
// 0..26 = the 27 neighbours).  Many tasks run concurrently (one scan has ~8 class tasks; a batch of scans
// fills the machine).
// The cluster -> instance-id order is the iteration order of the reference's
// std::unordered_map<int, vector<int>> (cluster_manager.hpp:395-418); it is reproduced
// on the host with the real container fed the labels in first-appearance order.
// Centroids are sequential float32 sums in ascending point order (get_json.cpp:266-274); they are summed per
// label on the GPU while the host works out the instance order.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <thread>
#include <unordered_map>

#include <cub/device/device_radix_sort.cuh>

#include "internal.cuh"

namespace sgtd {

constexpr int kS1Threads = 256;
constexpr int kMaxClass = 32;
constexpr int kEmptyVoxel = 0x7fffffff;
enum Policy : int { P_WHOLE = 0, P_GTINST = 1, P_DCVC = 2 };
enum Kind : int { K_NONE = 0, K_HEAD = 1, K_ALL = 2 };

struct Task {
  int scan, cls, policy, minSeg;
  int64_t pt0;      // first point of the scan in the batch arrays
  int npts_scan;    // points in the scan
  int npts;         // points of this class
  int64_t idx_off;  // into the per-point task arrays (cls_idx, slot, events, final, ...)
  int64_t tab_off;  // into the voxel-table arrays
  int tab_size;     // power of two
  int64_t lab_off;  // into the per-label arrays (parent, count, first): npts+1 entries (GTINST: 65536)
  int64_t bnd_off;  // into bounds (1024 per DCVC task)
};

struct TaskState {  // written by the device
  double minPitch, maxPitch, minPolar, maxPolar;
  int width, height, polarNum, nevents, labelCount, ndistinct;
  int nvox, ninvis;  // occupied voxels; seed events of invisible (top pitch layer) voxels
  int64_t pool_off;
  long long dbg_cyc[4];  // option s1_trace: cycles of the replay phases (build, translate, replay, write-back)
  int dbg_active, dbg_windows;
};

struct S1Buffers {
  const float4 *pts;
  const uint32_t *labels;
  const Task *tasks;
  TaskState *ts;
  int *cls_idx;      // original (scan-local) point index of each class point, ascending
  double *polar;     // 3 per point
  int *slot;         // voxel slot per point
  int4 *events;      // seed events in point order: (local rank, voxel slot, packed voxel coords, -)
  int2 *evc;         // replay form of the events: (table slot | which << 28 | visible << 30, local rank)
  int *pt_label;     // labels of points of invisible voxels
  int *final_label;  // final label per point
  int *parent, *count, *first;  // per label value
  int *t_key, *t_min1, *t_min2, *t_coord, *t_kind, *t_label;  // voxel table
  double *bounds;
  int *pool;         // (label, count, first) triples
  int *pool_of_label; // per label value: index of its pool entry (labels that own points)
  unsigned long long *pool_cursor;
};

// ---- K1: per-scan class histogram + "has a non-zero instance id" flags -------------
__global__ void k_s1_hist(const uint32_t *labels, const int64_t *scan_off, uint32_t *counts, uint32_t *nonzero,
                          uint32_t *bad) {
  __shared__ uint32_t s_cnt[kMaxClass], s_nz[kMaxClass];
  const int s = blockIdx.x;
  if (threadIdx.x < kMaxClass) { s_cnt[threadIdx.x] = 0; s_nz[threadIdx.x] = 0; }
  __syncthreads();
  const int64_t a = scan_off[s], b = scan_off[s + 1];
  for (int64_t i = a + blockIdx.y * blockDim.x + threadIdx.x; i < b; i += (int64_t)gridDim.y * blockDim.x) {
    const uint32_t l = labels[i];
    const uint32_t sem = l & 0xFFFFu;
    if (sem >= kMaxClass) { atomicOr(bad, 1u); continue; }
    atomicAdd(&s_cnt[sem], 1u);
    if (l >> 16) s_nz[sem] = 1u;
  }
  __syncthreads();
  if (threadIdx.x < kMaxClass) {
    if (s_cnt[threadIdx.x]) atomicAdd(&counts[s * kMaxClass + threadIdx.x], s_cnt[threadIdx.x]);
    if (s_nz[threadIdx.x]) atomicOr(&nonzero[s * kMaxClass + threadIdx.x], 1u);
  }
}

template <int kThreads = kS1Threads>
__device__ __forceinline__ int block_excl_scan(int v, int *s_warp, int &total) {
  // exclusive scan over the kThreads threads of the block (in thread order)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  __syncthreads();
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int base = 0; total = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) { if (w < warp) base += s_warp[w]; total += s_warp[w]; }
  return base + incl - v;
}

// ---- K2: ordered list of the scan's points that carry the task's class ----------------
__global__ void __launch_bounds__(kS1Threads) k_s1_gather(S1Buffers B) {
  __shared__ int s_warp[kS1Threads / 32];
  const Task t = B.tasks[blockIdx.x];
  constexpr int kPer = 8;  // consecutive points per thread and trip: one block scan per 2,048 points
  int base = 0;
  for (int i0 = 0; i0 < t.npts_scan; i0 += kS1Threads * kPer) {
    const int i = i0 + threadIdx.x * kPer;
    unsigned hits = 0;
    if (i + kPer <= t.npts_scan && (reinterpret_cast<uintptr_t>(B.labels + t.pt0 + i) & 15) == 0) {  // two 16-byte loads
      const uint4 a = *reinterpret_cast<const uint4 *>(B.labels + t.pt0 + i), b = *reinterpret_cast<const uint4 *>(B.labels + t.pt0 + i + 4);
      const uint32_t l[kPer] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int k = 0; k < kPer; ++k) hits |= (unsigned)((l[k] & 0xFFFFu) == (uint32_t)t.cls) << k;
    } else {
#pragma unroll
      for (int k = 0; k < kPer; ++k)
        if (i + k < t.npts_scan) hits |= (unsigned)((B.labels[t.pt0 + i + k] & 0xFFFFu) == (uint32_t)t.cls) << k;
    }
    int total;
    int pos = base + block_excl_scan(__popc(hits), s_warp, total);
#pragma unroll
    for (int k = 0; k < kPer; ++k)
      if ((hits >> k) & 1u) B.cls_idx[t.idx_off + pos++] = i + k;
    base += total;
  }
}

__device__ __forceinline__ uint32_t hash_i32(int k) {
  uint32_t x = (uint32_t)k;
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ int table_lookup(const int *t_key, int mask, int voxel) {
  uint32_t pos = hash_i32(voxel) & (uint32_t)mask;
  while (true) {
    const int k = __ldcg(t_key + pos);
    if (k == voxel) return (int)pos;
    if (k == kEmptyVoxel) return -1;
    pos = (pos + 1) & (uint32_t)mask;
  }
}

// ---- K3: polar transform, curved-voxel table, event list (one CTA per DCVC task) -------
// kThreads: 256 for tasks of up to kPrepSplit points, 1024 for larger ones (the loops are latency bound and the
// largest task of a batch is the kernel's critical path)
constexpr int kPrepSplit = 4096;
template <int kThreads>
__global__ void __launch_bounds__(kThreads) k_dcvc_prepare(S1Buffers B, double startR, double deltaR, double deltaP,
                                                            double deltaA, int *t_fe_all) {
  __shared__ int s_warp[kThreads / 32];
  __shared__ double s_red[4][kThreads / 32];
  __shared__ double s_mm[4];
  __shared__ int s_grid[3];
  const Task t = B.tasks[blockIdx.x];
  if (t.policy != P_DCVC || (kThreads == kS1Threads) != (t.npts <= kPrepSplit)) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double kPi = 3.14159265358979323846;  // M_PI
  double *polar = B.polar + 3 * t.idx_off;
  // 1. convert2polar (cluster_manager.hpp:172-206); running min/max start at 0,0,5,5 (:482-485)
  double mnP = 0.0, mxP = 0.0, mnR = 5.0, mxR = 5.0;
  for (int r = tid; r < t.npts; r += kThreads) {
    const float4 p = B.pts[t.pt0 + B.cls_idx[t.idx_off + r]];
    const double x = (double)p.x, y = (double)p.y, z = (double)p.z;
    const double rng = sqrt(__dadd_rn(__dmul_rn(x, x), __dadd_rn(__dmul_rn(y, y), __dmul_rn(z, z))));  // Eigen norm(): e0 + (e1 + e2)
    const double pitch = __dmul_rn(asin(z / rng), 180.0) / kPi;
    const double ang = atan2(y, x);
    const double az = ang > 0.0 ? __dmul_rn(ang, 180.0) / kPi : __dmul_rn(__dadd_rn(ang, __dmul_rn(2.0, kPi)), 180.0) / kPi;
    double o0 = 0.0, o1 = 0.0, o2 = 0.0;  // out-of-range points keep a zero polar record (SURVEY 8a note v)
    if (!(rng >= 120.0 || rng <= 0.5)) {
      mnP = pitch < mnP ? pitch : mnP; mxP = pitch > mxP ? pitch : mxP;
      mnR = rng < mnR ? rng : mnR; mxR = rng > mxR ? rng : mxR;
      o0 = rng; o1 = pitch; o2 = az;
    }
    polar[3 * r] = o0; polar[3 * r + 1] = o1; polar[3 * r + 2] = o2;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mnP = fmin(mnP, __shfl_xor_sync(0xffffffffu, mnP, o)); mxP = fmax(mxP, __shfl_xor_sync(0xffffffffu, mxP, o));
    mnR = fmin(mnR, __shfl_xor_sync(0xffffffffu, mnR, o)); mxR = fmax(mxR, __shfl_xor_sync(0xffffffffu, mxR, o));
  }
  if (lane == 0) { s_red[0][warp] = mnP; s_red[1][warp] = mxP; s_red[2][warp] = mnR; s_red[3][warp] = mxR; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kThreads / 32; ++w) {
      s_red[0][0] = fmin(s_red[0][0], s_red[0][w]); s_red[1][0] = fmax(s_red[1][0], s_red[1][w]);
      s_red[2][0] = fmin(s_red[2][0], s_red[2][w]); s_red[3][0] = fmax(s_red[3][0], s_red[3][w]);
    }
    const double minPitch = s_red[0][0], maxPitch = s_red[1][0], minPolar = s_red[2][0], maxPolar = s_red[3][0];
    // :212-220
    const int width = (int)(round(360.0 / deltaA) + 1.0);
    const int height = (int)(__dsub_rn(maxPitch, minPitch) / deltaP);
    double range = minPolar;
    int step = 1, polarNum = 0;
    double *bounds = B.bounds + t.bnd_off;
    while (range <= maxPolar && polarNum < 1024) {
      range = __dadd_rn(range, __dsub_rn(startR, __dmul_rn((double)step, deltaR)));
      bounds[polarNum] = range;
      ++polarNum; ++step;
    }
    s_mm[0] = minPitch; s_mm[1] = maxPitch; s_mm[2] = minPolar; s_mm[3] = maxPolar;
    s_grid[0] = width; s_grid[1] = height; s_grid[2] = polarNum;
    TaskState &ts = B.ts[blockIdx.x];
    ts.minPitch = minPitch; ts.maxPitch = maxPitch; ts.minPolar = minPolar; ts.maxPolar = maxPolar;
    ts.width = width; ts.height = height; ts.polarNum = polarNum;
  }
  __syncthreads();
  const double minPitch = s_mm[0];
  const int width = s_grid[0], height = s_grid[1], polarNum = s_grid[2];
  const double *bounds = B.bounds + t.bnd_off;
  int *t_key = B.t_key + t.tab_off, *t_min1 = B.t_min1 + t.tab_off, *t_min2 = B.t_min2 + t.tab_off;
  int *t_coord = B.t_coord + t.tab_off;
  const int mask = t.tab_size - 1;
  // 2. createHashTable (:224-252): voxel index of every point, table insert, lowest point per voxel
  for (int r = tid; r < t.npts; r += kThreads) {
    const double rng = polar[3 * r], pitch = polar[3 * r + 1], az = polar[3 * r + 2];
    // getPolarIndex (:259-264): first r with radius < bounds[r]; bounds are increasing here
    int lo = 0, hi = polarNum;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (rng < bounds[mid]) hi = mid; else lo = mid + 1; }
    const int polarIndex = lo < polarNum ? lo : polarNum - 1;
    const int pitchIndex = (int)round(__dsub_rn(pitch, minPitch) / deltaP);
    const int azIndex = (int)round(az / deltaA);
    const int voxel = (azIndex * (polarNum + 1) + polarIndex) + pitchIndex * (polarNum + 1) * (width + 1);
    uint32_t pos = hash_i32(voxel) & (uint32_t)mask;
    while (true) {
      const int old = atomicCAS(&t_key[pos], kEmptyVoxel, voxel);
      if (old == kEmptyVoxel) { t_coord[pos] = azIndex | (polarIndex << 10) | (pitchIndex << 21); break; }
      if (old == voxel) break;
      pos = (pos + 1) & (uint32_t)mask;
    }
    atomicMin(&t_min1[pos], r);
    B.slot[t.idx_off + r] = (int)pos;
  }
  __syncthreads();
  for (int r = tid; r < t.npts; r += kThreads) {
    const int pos = B.slot[t.idx_off + r];
    if (__ldcg(&t_min1[pos]) != r) atomicMin(&t_min2[pos], r);
  }
  __syncthreads();
  // 3. seed events in point order: two lowest points of a visible voxel, every point of an invisible one
  int base = 0;
  for (int r0 = 0; r0 < t.npts; r0 += kThreads) {
    const int r = r0 + tid;
    int ev = 0;
    if (r < t.npts) {
      const int pos = B.slot[t.idx_off + r];
      const int pitchIndex = __ldcg(&t_coord[pos]) >> 21;
      const bool visible = pitchIndex <= height;
      ev = !visible || __ldcg(&t_min1[pos]) == r || __ldcg(&t_min2[pos]) == r;
    }
    int total;
    const int p = block_excl_scan<kThreads>(ev, s_warp, total);
    if (ev) {
      const int pos = B.slot[t.idx_off + r];
      const int which = (__ldcg(&t_min1[pos]) == r ? 1 : 0) | (__ldcg(&t_min2[pos]) == r ? 2 : 0);
      B.events[t.idx_off + base + p] = make_int4(r, pos, __ldcg(&t_coord[pos]), which);
      if (which & 1) t_fe_all[t.tab_off + pos] = base + p;  // the voxel's name: index of its first seed event
    }
    base += total;
  }
  if (tid == 0) B.ts[blockIdx.x].nevents = base;
  // 4. voxel and invisible-seed counts (the replay picks its table class from them)
  int nv = 0, ni = 0;
  for (int sidx = tid; sidx < t.tab_size; sidx += kThreads) nv += __ldcg(&t_key[sidx]) != kEmptyVoxel;
  for (int r = tid; r < t.npts; r += kThreads) ni += (__ldcg(&t_coord[B.slot[t.idx_off + r]]) >> 21) > height;
  int tot_v, tot_i;
  block_excl_scan<kThreads>(nv, s_warp, tot_v);
  block_excl_scan<kThreads>(ni, s_warp, tot_i);
  if (tid == 0) { B.ts[blockIdx.x].nvox = tot_v; B.ts[blockIdx.x].ninvis = tot_i; }
}

__device__ __forceinline__ int uf_find_ro(const int *parent, int x) {
  while (true) { const int p = __ldcg(parent + x); if (p == x) return x; x = p; }
}

// ---- K4: replay of the seed events (DCVC, :272-355), one CTA per task -------------------------------
// The reference's walk is sequential over the seed events, but everything INSIDE one event is parallel:
//  * the curved-voxel table of the task lives in SHARED memory (open addressing; 4-byte word = packed
//    (azimuth, polar, pitch) index + 2-bit state, 2-byte label), rebuilt from the global table by the whole
//    CTA; tasks whose table does not fit use the global table (kSmem = false);
//  * warp 0 replays.  A window of 32 consecutive events is tested at once against the current state of
//    their own voxels: events whose point is already labelled (voxel ALL, or the other of its two seed
//    points) are skipped 32 at a time; the first still-unlabelled seed is the active one;
//  * for the active seed lanes 0..26 compute the 27 neighbour keys (searchKNN order, :365-385), look them
//    up and read their states in parallel.  The sequential walk (:317-346) collapses to warp primitives:
//    with r_k the union-find root of labelled neighbour k, `cur` only changes at the FIRST occurrence of
//    each distinct root (later occurrences are already merged), so the roots are chained
//    parent[r_f1] = r_f2, parent[r_f2] = r_f3, ... in first-occurrence order (the merge keeps the
//    neighbour's value, :323-327), an unlabelled neighbour takes the root of the last first-occurrence
//    before it, and the final `cur` is the root of the last one;
//  * the events stream through a small shared-memory ring that is refilled one chunk ahead.
// Results (voxel states, label forest) are written back to the global arrays k_s1_finish reads.
constexpr int kRpThreads = 128;
constexpr int kUfCap = 1024;     // union-find entries in shared memory; beyond: the task's global array
constexpr int kRing = 256;       // events in the ring (two chunks of 128)
constexpr int kRowRing = 128;    // neighbour rows kept ahead of the replaying warp
constexpr int kProducers = kRpThreads / 32 - 1;
constexpr uint32_t kVEmpty = 0xFFFFFFFFu;
constexpr uint32_t kCoordMask = 0x07FFFFFFu;  // az:9 | polar:10 | pitch:8

__device__ __forceinline__ uint32_t coord27(int az, int po, int pi) { return (uint32_t)az | ((uint32_t)po << 9) | ((uint32_t)pi << 19); }
__device__ __forceinline__ uint32_t hash_u32(uint32_t x) { x *= 0x9E3779B1u; return x ^ (x >> 15); }

template <bool kSmem>
struct VoxTable {
  // shared-memory form
  uint32_t *w; uint16_t *lab; uint32_t mask;
  // global form (the task's own table, keyed by the reference's voxel index)
  const int *t_key; int *t_kind, *t_label; int gmask; int polarNum, width;
  __device__ __forceinline__ int lookup(int az, int po, int pi) const {
    if (kSmem) {
      // linear probing in a sparsely filled table (load <= 0.3: most searches, hits or misses, end at the
      // first word).  The loop is kept convergent: a lane that is done idles until the slowest one is.
      const uint32_t key = coord27(az, po, pi);
      uint32_t pos = hash_u32(key) & mask;
      int found = -2;  // -2: still searching
      while (true) {
        if (found == -2) {
          const uint32_t v = w[pos];
          if (v == kVEmpty) found = -1;
          else if ((v & kCoordMask) == key) found = (int)pos;
          else pos = (pos + 1) & mask;
        }
        if (!__any_sync(__activemask(), found == -2)) return found;
      }
    } else {
      return table_lookup(t_key, gmask, (az * (polarNum + 1) + po) + pi * (polarNum + 1) * (width + 1));
    }
  }
  __device__ __forceinline__ void get(int slot, int &kind, int &label) const {
    if (kSmem) { kind = (int)(w[slot] >> 27) & 3; label = lab[slot]; }
    else { kind = __ldcg(t_kind + slot); label = __ldcg(t_label + slot); }
  }
  __device__ __forceinline__ int kind(int slot) const { return kSmem ? (int)(w[slot] >> 27) & 3 : __ldcg(t_kind + slot); }
  __device__ __forceinline__ void set(int slot, int kind, int label) {
    if (kSmem) { w[slot] = (w[slot] & kCoordMask) | ((uint32_t)kind << 27); lab[slot] = (uint16_t)label; }
    else { __stcg(t_kind + slot, kind); __stcg(t_label + slot, label); }
  }
  __device__ __forceinline__ void set_kind(int slot, int kind) {
    if (kSmem) w[slot] = (w[slot] & kCoordMask) | ((uint32_t)kind << 27);
    else __stcg(t_kind + slot, kind);
  }
};

struct UnionFind2 {
  int *sm, *gl;
  __device__ __forceinline__ int get(int x) const { return x < kUfCap ? sm[x] : __ldcg(gl + x); }
  __device__ __forceinline__ void set(int x, int v) { if (x < kUfCap) sm[x] = v; else __stcg(gl + x, v); }
  // read-mostly find with path halving; concurrent callers only ever move pointers towards the root
  __device__ __forceinline__ int find(int x) {
    while (true) {
      const int p = get(x);
      if (p == x) return x;
      const int g = get(p);
      if (g != p) set(x, g);
      x = g;
    }
  }
};

// nvox_lo < nvox <= nvox_hi selects the tasks of this launch (kSmem: table of `slots` entries in dynamic
// shared memory); seeds that can create more labels than a 16-bit label holds go to the global form.
template <bool kSmem>
__global__ void __launch_bounds__(kRpThreads) k_dcvc_replay(S1Buffers B, int slots, int nvox_lo, int nvox_hi, int use_rows, int cc_max_ev) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ int s_parent[kUfCap];
  __shared__ uint32_t s_stamp[kUfCap];  // per root: (event number << 5 | 31 - lowest lane that saw it in that event)
  __shared__ int2 s_ring[kRing];
  __shared__ int s_rows[kRowRing][32];      // neighbour slots of upcoming events, produced by warps 1..3
  __shared__ volatile int s_done[kRpThreads / 32];  // per producer warp: its events below this index are ready
  __shared__ volatile int s_cpos;           // the replaying warp's position in the event list
  const Task t = B.tasks[blockIdx.x];
  if (t.policy != P_DCVC) return;
  const TaskState ts = B.ts[blockIdx.x];
  const bool fits16 = ts.nvox + ts.ninvis < 65000;
  if (ts.nevents <= cc_max_ev) return;  // taken by k_dcvc_replay_cc
  if (kSmem) { if (!(ts.nvox > nvox_lo && ts.nvox <= nvox_hi && fits16)) return; }
  else if (ts.nvox <= nvox_lo && fits16) return;
  const int tid = threadIdx.x, lane = tid & 31;
  const int height = ts.height, polarNum = ts.polarNum, width = ts.width;
  const int nev = ts.nevents;
  VoxTable<kSmem> T{};
  T.w = reinterpret_cast<uint32_t *>(s_dyn); T.lab = reinterpret_cast<uint16_t *>(s_dyn + (size_t)slots * 4); T.mask = (uint32_t)slots - 1;
  T.t_key = B.t_key + t.tab_off; T.t_kind = B.t_kind + t.tab_off; T.t_label = B.t_label + t.tab_off; T.gmask = t.tab_size - 1;
  T.polarNum = polarNum; T.width = width;
  const int *g_coord = B.t_coord + t.tab_off;
  int4 *events = B.events + t.idx_off;
  int2 *evc = B.evc + t.idx_off;
  int *pt_label = B.pt_label + t.idx_off;
  const long long c0 = clock64();
  // ---- 1. the task's table in shared memory: one insert per voxel, from the event of its lowest point
  //         (event coord packs az | polar << 10 | pitch << 21)
  if (kSmem) {
    for (int i = tid; i < slots; i += kRpThreads) { T.w[i] = kVEmpty; T.lab[i] = 0; }
    __syncthreads();
#pragma unroll 4
    for (int e = tid; e < nev; e += kRpThreads) {
      const int4 ev = events[e];
      if (!(ev.w & 1)) continue;
      const int c = ev.z;
      const uint32_t key = coord27(c & 1023, (c >> 10) & 2047, c >> 21);  // kind NONE = 0 in bits 27..28
      uint32_t pos = hash_u32(key) & T.mask;
      while (atomicCAS(&T.w[pos], kVEmpty, key) != kVEmpty) pos = (pos + 1) & T.mask;
    }
    __syncthreads();
  }
  const long long c1 = clock64();
  for (int i = tid; i < kUfCap; i += kRpThreads) s_stamp[i] = 0;
  // ---- 2. replay form of the events: the slot of the seed's own voxel in the table that is replayed
#pragma unroll 4
  for (int e = tid; e < nev; e += kRpThreads) {
    const int4 ev = events[e];
    const int c = ev.z;
    const int slot = kSmem ? T.lookup(c & 1023, (c >> 10) & 2047, c >> 21) : ev.y;
    const int vis = (c >> 21) <= height;
    evc[e] = make_int2(slot | (ev.w << 28) | (vis << 30), ev.x);
  }
  if (tid < kRpThreads / 32) s_done[tid] = 0;
  if (tid == 0) s_cpos = 0;
  __threadfence_block();
  __syncthreads();
  // ---- 3. sequential replay by warp 0; warps 1..3 run ahead of it and look the 27 neighbour voxels of every
  //         event up (that part only depends on the table's keys, not on the labelling state)
  const long long c2 = clock64();
  int labelCount = 0, n_active = 0, n_windows = 0;
  // neighbour k = lane of the voxel in `slot` (searchKNN order, :365-385); -1 = skipped by the guards or absent
  auto neighbour = [&](int slot) -> int {
    int az, po, pi;
    if (kSmem) { const uint32_t c = T.w[slot] & kCoordMask; az = c & 511; po = (c >> 9) & 1023; pi = c >> 19; }
    else { const int c = __ldcg(g_coord + slot); az = c & 1023; po = (c >> 10) & 2047; pi = c >> 21; }
    int nb = -1;
    if (lane < 27) {
      const int z = pi - 1 + lane / 9, y = po - 1 + (lane / 3) % 3, x = az - 1 + lane % 3;
      if (!(z < 0 || z > height) && !(y < 0 || y > polarNum)) {
        int ax = x;
        if (ax < 0) ax = width - 1;
        if (ax > 300) ax = 300;
        if (y < polarNum) nb = T.lookup(ax, y, z);  // polar index polarNum is legal for the guard but never occupied
      }
    }
    return nb;
  };
  if (tid >= 32 && use_rows) {
    const int hw = tid >> 5;  // producer warp hw takes the events e with e % kProducers == hw - 1
    for (int e_base = hw - 1; e_base < nev; e_base += 32 * kProducers) {
      const int mine_e = e_base + lane * kProducers;
      const int2 mine = mine_e < nev ? evc[mine_e] : make_int2(0, 0);  // 32 events of this warp at a time
      for (int j = 0; j < 32; ++j) {
        const int e = e_base + j * kProducers;
        if (e >= nev) break;
        while (e >= s_cpos + kRowRing) __nanosleep(40);  // do not overwrite rows the replaying warp still needs
        const int slot = __shfl_sync(0xffffffffu, mine.x, j) & 0x0FFFFFFF;
        const int nb = neighbour(slot);
        s_rows[e & (kRowRing - 1)][lane] = nb;
        __syncwarp();
        if (lane == 0) { __threadfence_block(); s_done[hw] = e + 1; }
      }
    }
  }
  if (tid < 32) {
    UnionFind2 uf{s_parent, B.parent + t.lab_off};
    int loaded = 0, pre_base = 0;
    int2 pre[4];
    auto prefetch = [&](int base) {
      pre_base = base;
#pragma unroll
      for (int j = 0; j < 4; ++j) { const int i = base + 32 * j + lane; pre[j] = i < nev ? evc[i] : make_int2(0, 0); }
    };
    auto commit = [&]() {
#pragma unroll
      for (int j = 0; j < 4; ++j) s_ring[(pre_base + 32 * j + lane) & (kRing - 1)] = pre[j];
      loaded = min(nev, pre_base + 128);
      __syncwarp();
    };
    prefetch(0); commit(); prefetch(128);
    int e0 = 0;
    while (e0 < nev) {
      if (e0 + 32 > loaded && loaded < nev) { commit(); prefetch(loaded); }
      // ---- window: which of the next 32 seeds is still unlabelled?
      const int ei = e0 + lane;
      const bool valid = ei < nev;
      const int2 ew = valid ? s_ring[ei & (kRing - 1)] : make_int2(0, 0);
      const int oslot = ew.x & 0x0FFFFFFF, which = (ew.x >> 28) & 3, ovis = (ew.x >> 30) & 1;
      bool skip = !valid;
      if (valid && ovis) {  // (which: bit0 = lowest point of its voxel, bit1 = second lowest)
        const int kd = T.kind(oslot);
        skip = kd == K_ALL || (kd == K_NONE && !(which & 1)) || (kd == K_HEAD && !(which & 2));
      }
      const unsigned act = __ballot_sync(0xffffffffu, !skip);
      ++n_windows;
      if (act == 0) { e0 += 32; if (lane == 0) s_cpos = e0; continue; }
      ++n_active;
      const int a = __ffs(act) - 1;
      e0 += a + 1;
      const int v = __shfl_sync(0xffffffffu, oslot, a), vis = __shfl_sync(0xffffffffu, ovis, a);
      const int r = __shfl_sync(0xffffffffu, ew.y, a);
      // ---- the active seed: its 27 neighbour voxels (a visible voxel is its own neighbour 13), looked up
      //      ahead of time by a producer warp
      const int ea = e0 - 1;
      int nb;
      if (use_rows) {
        if (lane == 0) s_cpos = e0 - 1;
        while (s_done[ea % kProducers + 1] <= ea) { }
        __threadfence_block();
        nb = lane < 27 ? s_rows[ea & (kRowRing - 1)][lane] : -1;
      } else {
        nb = neighbour(v);
      }
      int kd = K_NONE, lb = -1;
      if (nb >= 0) T.get(nb, kd, lb);
      const bool labelled = nb >= 0 && kd != K_NONE;
      const unsigned Lm = __ballot_sync(0xffffffffu, labelled);

      int root = -1;
      if (labelled) root = uf.find(lb);
      __syncwarp();
      // first occurrence of each distinct root among the labelled lanes: the lowest lane wins a max-stamp
      // (stamps of earlier events are smaller); roots beyond the shared-memory range use the match instruction
      bool is_first = false;
      if (__any_sync(0xffffffffu, labelled && root >= kUfCap)) {
        unsigned grp = 0;
        if (labelled) grp = __match_any_sync(Lm, root);
        is_first = labelled && lane == __ffs(grp) - 1;
      } else {
        const uint32_t mine = ((uint32_t)n_active << 5) | (uint32_t)(31 - lane);
        if (labelled) atomicMax(&s_stamp[root], mine);
        __syncwarp();
        is_first = labelled && s_stamp[root] == mine;
      }
      const unsigned Fm = __ballot_sync(0xffffffffu, is_first);
      if (Fm == 0) {  // no labelled neighbour: new label for the seed and every neighbour (:340-346)
        const int L = ++labelCount;
        if (lane == 0) { uf.set(L, L); if (!vis) pt_label[r] = L; }
        if (nb >= 0) T.set(nb, K_ALL, L);
      } else {
        // chain the distinct roots in first-occurrence order: cur -> neigh (:323-327)
        const unsigned after = Fm & ~((2u << lane) - 1u);
        const int nroot = __shfl_sync(0xffffffffu, root, after ? __ffs(after) - 1 : lane);
        if (is_first && after) uf.set(root, nroot);
        const int cur = __shfl_sync(0xffffffffu, root, 31 - __clz(Fm));
        // an unlabelled neighbour takes the value `cur` has when the walk reaches it
        const unsigned before = Fm & ((1u << lane) - 1u);
        const int cprev = __shfl_sync(0xffffffffu, root, before ? 31 - __clz(before) : lane);
        const bool takes = nb >= 0 && kd == K_NONE && before;
        if (takes) T.set(nb, K_ALL, cprev);
        if (labelled && kd == K_HEAD) T.set_kind(nb, K_ALL);
        if (vis) {
          const unsigned self = __ballot_sync(0xffffffffu, nb == v && (takes || labelled));
          // own voxel passed while cur == -1: only the seed point itself is labelled (:330-331)
          if (!self && lane == 0) T.set(v, K_HEAD, cur);
        } else if (lane == 0) {
          pt_label[r] = cur;
        }
      }
      // Voxel states (global form) and label-forest entries beyond the shared-memory range live in global
      // memory, written with plain stores by some lanes and read with ld.cg by others in the next event:
      // __syncwarp alone did not order them on B200 (observed as wrong merges once the replaying warp got
      // faster), a block-level fence does.
      if (!kSmem || labelCount >= kUfCap) __threadfence_block();
      __syncwarp();
    }
    // publish the shared-memory part of the label forest for k_s1_finish
    for (int i = 1 + lane; i <= labelCount && i < kUfCap; i += 32) uf.gl[i] = s_parent[i];
    if (lane == 0) { B.ts[blockIdx.x].labelCount = labelCount; B.ts[blockIdx.x].dbg_active = n_active; B.ts[blockIdx.x].dbg_windows = n_windows; }
  }
  __syncthreads();
  const long long c3 = clock64();
  // ---- 4. voxel states back to the global table (one write per voxel, through its lowest point's event)
  if (kSmem) {
#pragma unroll 4
    for (int e = tid; e < nev; e += kRpThreads) {
      const int4 ev = events[e];
      if (!(ev.w & 1)) continue;
      const int slot = evc[e].x & 0x0FFFFFFF;
      int kd, lb;
      T.get(slot, kd, lb);
      T.t_kind[ev.y] = kd; T.t_label[ev.y] = kd == K_NONE ? -1 : lb;
    }
  }
  if (tid == 0) {
    long long *d = B.ts[blockIdx.x].dbg_cyc;
    d[0] = c1 - c0; d[1] = c2 - c1; d[2] = c3 - c2; d[3] = clock64() - c3;
  }
}

// ---- K4b: component-parallel replay ---------------------------------------------------------------
// Seed events only read and write the states of the seed's <= 27 neighbour voxels (and the labels living on
// them), so events of different connected components of the voxel-neighbour graph commute; the only thing
// they share is the counter that names new labels.  Two kernels:
//   k_dcvc_rows   (fully parallel, all tasks) names every voxel after its FIRST seed event and writes, per
//                 voxel, the row of its 27 neighbours (searchKNN order, :365-385) as such names -- the only
//                 place the hash table is read;
//   k_dcvc_replay_cc (one CTA of 8 warps per task; per-voxel state = 1-byte kind + 2-byte label in shared
//                 memory, indexed by the voxel's name: no table in shared memory)
//     1. unions every voxel with its row (16-bit union-find, atomicCAS hooks the larger name under the
//        smaller): the components of the neighbour graph;
//     2. hands the components to the warps (largest first, then greedily to the lightest warp, weight =
//        events) and tags every event with its owner;
//     3. lets all warps replay concurrently: a warp walks the event list with the same 32-wide windows as the
//        sequential form, events of other warps count as skipped.  A new label is provisionally NAMED after
//        the event that creates it (event index + 1); merges (parent[cur] = neigh) are structural;
//     4. renames: the reference's label value is 1 + the number of label-creating events before it -- a
//        prefix sum over the creation bitmap -- applied while the voxel states, the label forest and the
//        labels of invisible seeds are written back.
// Taken by tasks with at most kCcMaxEv events (three shared-memory classes by event count); the rest go
// through the sequential forms above.
constexpr int kCcThreads = 256;
constexpr int kCcWarps = kCcThreads / 32;
constexpr int kCcMaxEv = 16384;
constexpr int kCcRing = 256;
constexpr uint16_t kNoVox = 0xFFFFu;

__global__ void __launch_bounds__(kS1Threads) k_dcvc_rows(S1Buffers B, const int *t_fe_all, uint16_t *rows_all) {
  const Task t = B.tasks[blockIdx.x];
  if (t.policy != P_DCVC) return;
  const TaskState ts = B.ts[blockIdx.x];
  const int nev = ts.nevents;
  if (nev > kCcMaxEv) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int height = ts.height, polarNum = ts.polarNum, width = ts.width;
  const int *t_key = B.t_key + t.tab_off, *t_fe = t_fe_all + t.tab_off;
  const int gmask = t.tab_size - 1;
  const int4 *events = B.events + t.idx_off;
  int2 *evc = B.evc + t.idx_off;
  uint16_t *rows = rows_all + (size_t)t.idx_off * 32;
  // grid (task, chunk of kS1Threads events): big tasks are spread over many CTAs
  for (int e0 = blockIdx.y * kS1Threads + wid * 32; e0 < nev; e0 += gridDim.y * kS1Threads) {
    const int e = e0 + lane;
    int4 ev = make_int4(0, 0, 0, 0);
    if (e < nev) {
      ev = events[e];
      const int vis = (ev.z >> 21) <= height;
      evc[e] = make_int2(__ldcg(t_fe + ev.y) | (ev.w << 28) | (vis << 30), ev.x);
    }
    unsigned firsts = __ballot_sync(0xffffffffu, e < nev && (ev.w & 1));
    while (firsts) {
      const int j = __ffs(firsts) - 1;
      firsts &= firsts - 1;
      const int c = __shfl_sync(0xffffffffu, ev.z, j);
      const int az = c & 1023, po = (c >> 10) & 2047, pi = c >> 21;
      uint16_t out = kNoVox;
      if (lane < 27) {
        const int z = pi - 1 + lane / 9, y = po - 1 + (lane / 3) % 3, x = az - 1 + lane % 3;
        if (!(z < 0 || z > height) && !(y < 0 || y > polarNum)) {
          int ax = x;
          if (ax < 0) ax = width - 1;
          if (ax > 300) ax = 300;
          if (y < polarNum) {  // polar index polarNum is legal for the guard but never occupied
            const int pos = table_lookup(t_key, gmask, (ax * (polarNum + 1) + y) + z * (polarNum + 1) * (width + 1));
            if (pos >= 0) out = (uint16_t)__ldcg(t_fe + pos);
          }
        }
      }
      rows[(size_t)(e0 + j) * 32 + lane] = out;
    }
  }
}

__device__ __forceinline__ int uf16_find(volatile uint16_t *par, int x) {
  while (true) {
    const int p = par[x];
    if (p == x) return x;
    const int g = par[p];
    if (g != p) par[x] = (uint16_t)g;  // path halving; concurrent callers only move pointers towards the root
    x = g;
  }
}
__device__ __forceinline__ void uf16_unite(uint16_t *par, int a, int b) {
  while (true) {
    a = uf16_find(par, a); b = uf16_find(par, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }
    if (atomicCAS(reinterpret_cast<unsigned short *>(par + a), (unsigned short)a, (unsigned short)b) == (unsigned short)a) return;
  }
}

// cap: events the shared-memory arrays hold; tasks with nev_lo < nevents <= cap are taken
__global__ void __launch_bounds__(kCcThreads) k_dcvc_replay_cc(S1Buffers B, const uint16_t *rows_all, int cap, int nev_lo) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ uint32_t s_new[kCcMaxEv / 32 + 1];  // bit e: event e created a label
  __shared__ uint32_t s_pre[kCcMaxEv / 32 + 1];  // labels created before word w
  __shared__ int s_tot[4];
  __shared__ int s_wload[kCcWarps];
  const Task t = B.tasks[blockIdx.x];
  if (t.policy != P_DCVC) return;
  const TaskState ts = B.ts[blockIdx.x];
  const int nev = ts.nevents;
  if (!(nev > nev_lo && nev <= cap)) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // per voxel (named after its first event): label, kind; per event / label: union-find, root
  uint16_t *s_lab = reinterpret_cast<uint16_t *>(s_dyn);
  uint16_t *s_a = s_lab + cap;             // union-find of the components, then of the labels (index = label <= cap)
  uint16_t *s_b = s_a + (cap + 2);         // component root of every event
  int2 *s_ring = reinterpret_cast<int2 *>(s_b + cap + 2);  // (6 cap + 8 bytes: 8-byte aligned for cap % 4 == 0)
  uint8_t *s_kind = reinterpret_cast<uint8_t *>(s_ring + kCcWarps * kCcRing);
  const int4 *events = B.events + t.idx_off;
  int2 *evc = B.evc + t.idx_off;
  int *pt_label = B.pt_label + t.idx_off;
  const uint16_t *rows = rows_all + (size_t)t.idx_off * 32;
  const long long c0 = clock64();
  for (int e = tid; e < nev; e += kCcThreads) { s_a[e] = (uint16_t)e; s_kind[e] = K_NONE; s_lab[e] = 0; }
  const int nwords = (nev + 31) >> 5;
  for (int i = tid; i <= nwords; i += kCcThreads) s_new[i] = 0;
  if (tid < 4) s_tot[tid] = 0;
  if (tid < kCcWarps) s_wload[tid] = 0;
  __syncthreads();
  // ---- 1. components: every voxel united with its neighbour row.  Warp-cooperative: the roots of the voxel
  //         and of its neighbours are found in parallel, the smallest is the target and every other distinct
  //         root is hooked under it with one CAS (a lost race falls back to the generic unite); the next
  //         voxel's row is in flight meanwhile.
  for (int e0 = wid * 32; e0 < nev; e0 += kCcThreads) {
    const int e = e0 + lane;
    const int x = e < nev ? evc[e].x : 0;
    unsigned firsts = __ballot_sync(0xffffffffu, e < nev && ((x >> 28) & 1));
    while (firsts) {
      // up to kDepth voxels per trip: all their rows are requested before the first is used
      constexpr int kDepth = 8;
      int jj[kDepth];
      uint16_t nbs[kDepth];
#pragma unroll
      for (int u = 0; u < kDepth; ++u) {
        jj[u] = firsts ? __ffs(firsts) - 1 : -1;
        firsts &= firsts - 1;
        nbs[u] = jj[u] >= 0 ? rows[(size_t)(e0 + jj[u]) * 32 + lane] : kNoVox;
      }
#pragma unroll
      for (int u = 0; u < kDepth; ++u) {
        if (jj[u] < 0) break;
        const int mine = nbs[u] != kNoVox ? (int)nbs[u] : e0 + jj[u];  // lanes without a neighbour stand in for the voxel itself
        const int rk = uf16_find(s_a, mine);
        const int m = (int)__reduce_min_sync(0xffffffffu, (unsigned)rk);
        if (rk != m && atomicCAS(reinterpret_cast<unsigned short *>(s_a + rk), (unsigned short)rk, (unsigned short)m) != (unsigned short)rk)
          uf16_unite(s_a, rk, m);
      }
    }
  }
  __syncthreads();
  const long long c_union = clock64();
  // ---- 2. root and weight of every component, owners
  for (int e = tid; e < nev; e += kCcThreads) s_b[e] = (uint16_t)uf16_find(s_a, evc[e].x & 0xFFFF);
  __syncthreads();
  uint32_t *s_a32 = reinterpret_cast<uint32_t *>(s_a);
  for (int i = tid; i < (nev + 2) / 2; i += kCcThreads) s_a32[i] = 0;
  __syncthreads();
  for (int e = tid; e < nev; e += kCcThreads) { const uint32_t r = s_b[e]; atomicAdd(&s_a32[r >> 1], 1u << (16 * (r & 1u))); }  // <= 16384 per half
  __syncthreads();
  const long long c_wgt = clock64();
  if (wid == 0) {
    int load[kCcWarps];
#pragma unroll
    for (int w = 0; w < kCcWarps; ++w) load[w] = 0;
    // two passes: components with at least 1/16 of the events first (they decide the balance), then the rest
    const int big = max(nev / 16, 1);
    for (int pass = 0; pass < 2; ++pass)
      for (int e0 = 0; e0 < nev; e0 += 32) {
        const int e = e0 + lane;
        const bool is_root = e < nev && s_b[e] == (uint16_t)e;
        const int mine = is_root ? (int)s_a[e] : 0;
        unsigned roots = __ballot_sync(0xffffffffu, is_root && ((pass == 0) == (mine >= big)));
        while (roots) {
          const int j = __ffs(roots) - 1;
          roots &= roots - 1;
          const int wgt = __shfl_sync(0xffffffffu, mine, j);
          int best = 0;
#pragma unroll
          for (int w = 1; w < kCcWarps; ++w) if (load[w] < load[best]) best = w;
#pragma unroll
          for (int w = 0; w < kCcWarps; ++w) if (w == best) load[w] += wgt;
          if (lane == 0) s_lab[e0 + j] = (uint16_t)best;  // (s_lab is free until the replay starts)
        }
      }
  }
  __syncthreads();
#pragma unroll 4
  for (int e = tid; e < nev; e += kCcThreads) { int2 v = evc[e]; v.x |= (int)s_lab[s_b[e]] << 16; evc[e] = v; }
  __syncthreads();
  for (int e = tid; e < nev; e += kCcThreads) s_lab[e] = 0;
  __threadfence_block();
  __syncthreads();
  const long long c1 = clock64();
  // ---- 3. replay: every warp walks the whole event list and takes the seeds of its own components
  {
    uint16_t *lpar = s_a;  // label forest, indexed by the provisional label (creating event + 1)
    int2 *ring = s_ring + wid * kCcRing;
    auto lfind = [&](int x) -> int {
      while (true) {
        const int p = lpar[x];
        if (p == x) return x;
        const int g = lpar[p];
        if (g != p) lpar[x] = (uint16_t)g;
        x = g;
      }
    };
    auto set_vox = [&](int vx, int kind, int label) { s_kind[vx] = (uint8_t)kind; s_lab[vx] = (uint16_t)label; };
    int n_active = 0, n_windows = 0, n_labels = 0;
    int loaded = 0, pre_base = 0;
    int2 pre[4];
    auto prefetch = [&](int base) {
      pre_base = base;
#pragma unroll
      for (int j = 0; j < 4; ++j) { const int i = base + 32 * j + lane; pre[j] = i < nev ? evc[i] : make_int2(0, 0); }
    };
    auto commit = [&]() {
      __syncwarp();  // the slots overwritten here were read (windows 256 events back) by every lane
#pragma unroll
      for (int j = 0; j < 4; ++j) ring[(pre_base + 32 * j + lane) & (kCcRing - 1)] = pre[j];
      loaded = min(nev, pre_base + 128);
      __syncwarp();
    };
    prefetch(0); commit(); prefetch(128);
    int e0 = 0;
    while (e0 < nev) {
      if (e0 + 32 > loaded && loaded < nev) { commit(); prefetch(loaded); }
      const int ei = e0 + lane;
      const bool valid = ei < nev;
      const int2 ew = valid ? ring[ei & (kCcRing - 1)] : make_int2(0, 0);
      const int ovox = ew.x & 0xFFFF, owner = (ew.x >> 16) & 7, which = (ew.x >> 28) & 3, ovis = (ew.x >> 30) & 1;
      bool skip = !valid || owner != wid;
      if (!skip && ovis) {
        const int kd = s_kind[ovox];
        skip = kd == K_ALL || (kd == K_NONE && !(which & 1)) || (kd == K_HEAD && !(which & 2));
      }
      const unsigned act = __ballot_sync(0xffffffffu, !skip);
      ++n_windows;
      if (act == 0) { e0 += 32; continue; }
      ++n_active;
      const int a = __ffs(act) - 1;
      e0 += a + 1;
      const int v = __shfl_sync(0xffffffffu, ovox, a), vis = __shfl_sync(0xffffffffu, ovis, a);
      const int r = __shfl_sync(0xffffffffu, ew.y, a);
      const int ea = e0 - 1;
      int nb = -1;
      if (lane < 27) { const uint16_t x = rows[(size_t)v * 32 + lane]; nb = x == kNoVox ? -1 : (int)x; }
      {
        // rows of the seeds this warp may take next (its events among the 32 after this one): into L1 while
        // this seed is walked
        const int en = e0 + lane;
        if (en < min(loaded, nev)) {
          const int xn = ring[en & (kCcRing - 1)].x;
          if (((xn >> 16) & 7) == wid) asm volatile("prefetch.global.L1 [%0];" ::"l"(rows + (size_t)(xn & 0xFFFF) * 32));
        }
      }
      int kd = K_NONE, lb = -1;
      if (nb >= 0) { kd = s_kind[nb]; lb = s_lab[nb]; }
      const bool labelled = nb >= 0 && kd != K_NONE;
      const unsigned Lm = __ballot_sync(0xffffffffu, labelled);
      int root = -1;
      if (labelled) root = lfind(lb);
      __syncwarp();
      bool is_first = false;
      if (labelled) { const unsigned grp = __match_any_sync(Lm, root); is_first = lane == __ffs(grp) - 1; }
      const unsigned Fm = __ballot_sync(0xffffffffu, is_first);
      if (Fm == 0) {  // no labelled neighbour: new label for the seed and every neighbour (:340-346)
        const int L = ea + 1;
        ++n_labels;
        if (lane == 0) { lpar[L] = (uint16_t)L; atomicOr(&s_new[ea >> 5], 1u << (ea & 31)); if (!vis) pt_label[r] = L; }
        if (nb >= 0) set_vox(nb, K_ALL, L);
      } else {
        // chain the distinct roots in first-occurrence order: cur -> neigh (:323-327)
        const unsigned after = Fm & ~((2u << lane) - 1u);
        const int nroot = __shfl_sync(0xffffffffu, root, after ? __ffs(after) - 1 : lane);
        if (is_first && after) lpar[root] = (uint16_t)nroot;
        const int cur = __shfl_sync(0xffffffffu, root, 31 - __clz(Fm));
        // an unlabelled neighbour takes the value `cur` has when the walk reaches it
        const unsigned before = Fm & ((1u << lane) - 1u);
        const int cprev = __shfl_sync(0xffffffffu, root, before ? 31 - __clz(before) : lane);
        const bool takes = nb >= 0 && kd == K_NONE && before;
        if (takes) set_vox(nb, K_ALL, cprev);
        if (labelled && kd == K_HEAD) s_kind[nb] = K_ALL;
        if (vis) {
          const unsigned self = __ballot_sync(0xffffffffu, nb == v && (takes || labelled));
          // own voxel passed while cur == -1: only the seed point itself is labelled (:330-331)
          if (!self && lane == 0) set_vox(v, K_HEAD, cur);
        } else if (lane == 0) {
          pt_label[r] = cur;
        }
      }
      __syncwarp();
    }
    if (lane == 0) { atomicAdd(&s_tot[0], n_active); atomicAdd(&s_tot[1], n_windows); atomicAdd(&s_tot[2], n_labels); s_wload[wid] = n_active; }
  }
  __threadfence_block();
  __syncthreads();
  const long long c2 = clock64();
  // ---- 4. the reference's label values, write-back
  if (wid == 0) {
    int run = 0;
    for (int w0 = 0; w0 <= nwords; w0 += 32) {
      const int w = w0 + lane;
      const int cnt = (w <= nwords) ? __popc(s_new[w]) : 0;
      int inc = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
      if (w <= nwords) s_pre[w] = (uint32_t)(run + inc - cnt);
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  __syncthreads();
  auto final_label = [&](int L) -> int {  // provisional (creating event + 1) -> 1 + number of creating events before it
    const int e = L - 1;
    return (int)s_pre[e >> 5] + __popc(s_new[e >> 5] & ((1u << (e & 31)) - 1u)) + 1;
  };
  int *g_kind = B.t_kind + t.tab_off, *g_label = B.t_label + t.tab_off, *g_parent = B.parent + t.lab_off;
#pragma unroll 2
  for (int e = tid; e < nev; e += kCcThreads) {
    const int4 ev = events[e];
    const int2 ec = evc[e];
    if (ev.w & 1) {
      const int kd = s_kind[e];
      g_kind[ev.y] = kd; g_label[ev.y] = kd == K_NONE ? -1 : final_label((int)s_lab[e]);
    }
    if (!((ec.x >> 30) & 1)) pt_label[ec.y] = final_label(pt_label[ec.y]);
    if ((s_new[e >> 5] >> (e & 31)) & 1u) g_parent[final_label(e + 1)] = final_label((int)s_a[e + 1]);
  }
  if (tid == 0) {
    TaskState &o = B.ts[blockIdx.x];
    int mx = 0;
    for (int w = 0; w < kCcWarps; ++w) mx = max(mx, s_wload[w]);
    o.labelCount = s_tot[2]; o.dbg_active = s_tot[0]; o.dbg_windows = mx;  // (windows field: seeds of the busiest warp)
    // (trace fields: union | weights + owners + tagging | replay + renaming + write-back | of which owners + tagging)
    o.dbg_cyc[0] = c_union - c0; o.dbg_cyc[1] = c1 - c_union; o.dbg_cyc[2] = clock64() - c1; o.dbg_cyc[3] = c1 - c_wgt;
    (void)c2;
  }
}
static size_t cc_smem(int cap) { return (size_t)cap * 7 + 16 + (size_t)kCcWarps * kCcRing * sizeof(int2); }

// ---- K5: final label per point, per-label size and first point, compact label list -------
// (kThreads: 256 for tasks of up to kPrepSplit points, 1024 above -- as k_dcvc_prepare)
template <int kThreads>
__global__ void __launch_bounds__(kThreads) k_s1_finish(S1Buffers B) {
  __shared__ int s_warp[kThreads / 32];
  __shared__ long long s_pool;
  const Task t = B.tasks[blockIdx.x];
  if ((kThreads == kS1Threads) != (t.npts <= kPrepSplit)) return;
  const int tid = threadIdx.x;
  int *count = B.count + t.lab_off, *first = B.first + t.lab_off;
  int nlab = 0;
  if (t.policy == P_DCVC) {
    const TaskState ts = B.ts[blockIdx.x];
    const int *parent = B.parent + t.lab_off;
    const int *t_coord = B.t_coord + t.tab_off, *t_label = B.t_label + t.tab_off;
    for (int r = tid; r < t.npts; r += kThreads) {
      const int v = B.slot[t.idx_off + r];
      const bool vis = (__ldcg(t_coord + v) >> 21) <= ts.height;
      const int raw = vis ? __ldcg(t_label + v) : B.pt_label[t.idx_off + r];
      const int lab = uf_find_ro(parent, raw);
      B.final_label[t.idx_off + r] = lab;
      atomicAdd(&count[lab], 1);
      atomicMin(&first[lab], r);
    }
    nlab = ts.labelCount + 1;
  } else if (t.policy == P_GTINST) {
    for (int r = tid; r < t.npts; r += kThreads) {
      const int lab = (int)(B.labels[t.pt0 + B.cls_idx[t.idx_off + r]] >> 16);
      B.final_label[t.idx_off + r] = lab;
      atomicAdd(&count[lab], 1);
      atomicMin(&first[lab], r);
    }
    nlab = 65536;
  } else {
    for (int r = tid; r < t.npts; r += kThreads) B.final_label[t.idx_off + r] = 0;
    if (tid == 0) { count[0] = t.npts; first[0] = 0; }
    nlab = 1;
  }
  __syncthreads();
  // distinct labels -> pool (label, count, first); order is fixed on the host
  int mine = 0;
  for (int l = tid; l < nlab; l += kThreads) mine += __ldcg(&count[l]) > 0;
  int total;
  int pos = block_excl_scan<kThreads>(mine, s_warp, total);
  if (tid == 0) {
    s_pool = (long long)atomicAdd(B.pool_cursor, (unsigned long long)total);
    B.ts[blockIdx.x].pool_off = s_pool;
    B.ts[blockIdx.x].ndistinct = total;
  }
  __syncthreads();
  for (int l = tid; l < nlab; l += kThreads) {
    const int c = __ldcg(&count[l]);
    if (c > 0) {
      int *o = B.pool + 3 * (s_pool + pos);
      o[0] = l; o[1] = c; o[2] = __ldcg(&first[l]);
      B.pool_of_label[t.lab_off + l] = (int)(s_pool + pos);
      ++pos;
    }
  }
}

struct InstRec {  // one per instance, host-planned
  int task, label, inst_id, node_slot;  // node_slot: index into the node output or -1
  uint32_t node_label;
  int pool_idx;                         // pool entry of the label (its centroid sits at that index)
};

// The centroid of a node instance only depends on the label's points, not on the instance order the host
// works out from the pool (unordered_map replay): K6a-K6c run on the GPU WHILE the host orders.
// ---- K6a: per point (sort key, position): key = pool index of its label if the label will become a node
//           instance (node class and large enough: get_json.cpp:146,228 / minSeg), else all-ones.
__global__ void __launch_bounds__(kS1Threads) k_s1_lkey(S1Buffers B, uint32_t *skey, uint32_t *sval) {
  const Task t = B.tasks[blockIdx.x];
  const bool node = t.cls >= 10 && t.cls <= 18;  // node_map(cls) in 3..11
  const int *count = B.count + t.lab_off;
  for (int r = threadIdx.x; r < t.npts; r += kS1Threads) {
    const int lab = B.final_label[t.idx_off + r];
    uint32_t key = 0xFFFFFFFFu;
    if (node) {
      const int c = count[lab];
      const bool ok = t.policy == P_DCVC ? c >= t.minSeg : (t.policy == P_GTINST ? c > 20 : true);
      if (ok) key = (uint32_t)B.pool_of_label[t.lab_off + lab];
    }
    skey[t.idx_off + r] = key;
    sval[t.idx_off + r] = (uint32_t)(t.pt0 + B.cls_idx[t.idx_off + r]);
  }
}

// ---- K6b (after a stable sort by key): centroid = sequential float32 sum in ascending point order
// (get_json.cpp:266-274).  One warp per pool entry that owns sorted points (found by binary search):
// 32 x kDepth points are fetched per trip (coalesced list, gathered coordinates) and added in lane order, so
// the rounding sequence is the reference's.
__global__ void __launch_bounds__(128) k_s1_lcentroid(S1Buffers B, const uint32_t *sorted_key, const uint32_t *sorted_pos,
                                                      int64_t n_idx, float4 *cent) {
  const int lane = threadIdx.x & 31;
  const long long npool = (long long)*B.pool_cursor;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < npool; i += nwarps) {
    const int64_t n = B.pool[3 * i + 1];
    int64_t lo = 0, hi = n_idx;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (sorted_key[mid] < (uint32_t)i) lo = mid + 1; else hi = mid; }
    if (lo >= n_idx || sorted_key[lo] != (uint32_t)i) continue;  // not a node instance
    const int64_t s0 = lo;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    constexpr int kDepth = 4;  // 128 points per trip; the next trip's two dependent loads are in flight while this one is summed
    float4 q[kDepth];
    auto fetch = [&](int64_t j0) {
#pragma unroll
      for (int u = 0; u < kDepth; ++u) {
        const int64_t j = j0 + 32 * u + lane;
        q[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < n) q[u] = B.pts[sorted_pos[s0 + j]];
      }
    };
    fetch(0);
    for (int64_t j0 = 0; j0 < n; j0 += 32 * kDepth) {
      float4 p[kDepth];
#pragma unroll
      for (int u = 0; u < kDepth; ++u) p[u] = q[u];
      if (j0 + 32 * kDepth < n) fetch(j0 + 32 * kDepth);
#pragma unroll
      for (int u = 0; u < kDepth; ++u) {
        const int m = (int)max((int64_t)0, min((int64_t)32, n - j0 - 32 * u));
        if (m == 32) {
          // full tile: unrolled, so that the broadcasts pipeline and only the three add chains are serial
#pragma unroll
          for (int l = 0; l < 32; ++l) {
            cx = __fadd_rn(cx, __shfl_sync(0xffffffffu, p[u].x, l));
            cy = __fadd_rn(cy, __shfl_sync(0xffffffffu, p[u].y, l));
            cz = __fadd_rn(cz, __shfl_sync(0xffffffffu, p[u].z, l));
          }
        } else {
          for (int l = 0; l < m; ++l) {
            cx = __fadd_rn(cx, __shfl_sync(0xffffffffu, p[u].x, l));
            cy = __fadd_rn(cy, __shfl_sync(0xffffffffu, p[u].y, l));
            cz = __fadd_rn(cz, __shfl_sync(0xffffffffu, p[u].z, l));
          }
        }
      }
    }
    if (lane == 0) {
      const float cnt = (float)n;
      cent[i] = make_float4(__fdiv_rn(cx, cnt), __fdiv_rn(cy, cnt), __fdiv_rn(cz, cnt), 0.f);
    }
  }
}

// ---- K7 (after the host has ordered the instances): label -> instance id, node records, membership
__global__ void k_s1_scatter_map(S1Buffers B, const InstRec *inst, int ninst, int *inst_of_label, const float4 *cent, sgtd_node *nodes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ninst) return;
  const InstRec ir = inst[i];
  const Task t = B.tasks[ir.task];
  inst_of_label[t.lab_off + ir.label] = ir.inst_id;
  if (ir.node_slot >= 0) {
    const float4 c = cent[ir.pool_idx];
    sgtd_node nd;
    nd.x = c.x; nd.y = c.y; nd.z = c.z; nd.label = ir.node_label;
    nodes[ir.node_slot] = nd;
  }
}
__global__ void __launch_bounds__(kS1Threads) k_s1_assign(S1Buffers B, const int *inst_of_label, int32_t *point_instance) {
  const Task t = B.tasks[blockIdx.x];
  for (int r = threadIdx.x; r < t.npts; r += kS1Threads) {
    const int id = inst_of_label[t.lab_off + B.final_label[t.idx_off + r]];
    if (id >= 0) point_instance[t.pt0 + B.cls_idx[t.idx_off + r]] = id;
  }
}

__global__ void k_fill_i32(int *p, int64_t n, int v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

static int node_map(int c) {  // R/src/get_json.cpp:10-12 ; -1 = key absent (label 9 / unknown: skipped)
  switch (c) {
    case 10: return 3; case 11: return 4; case 12: return 5; case 13: return 6; case 14: return 7;
    case 15: return 8; case 16: return 9; case 17: return 10; case 18: return 11;
    case 0: case 1: case 2: case 3: case 4: case 5: case 6: case 7: case 8: case 19: return 0;
    default: return -1;
  }
}

struct S1Pool {
  DevBuf<int64_t> d_off; DevBuf<uint32_t> d_cnt; DevBuf<Task> d_tasks; DevBuf<TaskState> d_ts;
  DevBuf<int> d_pp, d_tab, d_lab, d_pool, d_map; DevBuf<double> d_polar, d_bounds; DevBuf<unsigned long long> d_cur;
  DevBuf<InstRec> d_inst; DevBuf<sgtd_node> d_nodes;
  DevBuf<uint16_t> d_rows; DevBuf<int> d_map2; DevBuf<uint32_t> d_sort; DevBuf<int64_t> d_seg; DevBuf<unsigned char> d_cub;
  DevBuf<float4> d_cent; DevBuf<float4> in_pts; DevBuf<uint32_t> in_lab; DevBuf<int32_t> out_pi;  // staging of host inputs / outputs
  cudaStream_t side[4] = {}; cudaEvent_t ev[5] = {}; bool have_streams = false;  // concurrent replay classes
  ~S1Pool() {
    if (have_streams) { for (auto x : side) cudaStreamDestroy(x); for (auto x : ev) cudaEventDestroy(x); }
    d_off.release(); d_cnt.release(); d_tasks.release(); d_ts.release(); d_pp.release(); d_tab.release(); d_lab.release();
    d_pool.release(); d_map.release(); d_polar.release(); d_bounds.release(); d_cur.release(); d_inst.release();
    d_nodes.release(); in_pts.release(); in_lab.release(); out_pi.release();
    d_cent.release(); d_rows.release(); d_map2.release(); d_sort.release(); d_seg.release(); d_cub.release();
  }
};
static void s1_pool_free(void *p) { delete static_cast<S1Pool *>(p); }
static S1Pool &s1_pool(sgtd_handle *h) {
  if (!h->s1pool) { h->s1pool = new S1Pool(); h->s1pool_free = s1_pool_free; }
  return *static_cast<S1Pool *>(h->s1pool);
}

#define S1_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { rc = sgtd_fail(h, SGTD_E_CUDA, #expr, __FILE__, __LINE__, _e); goto done; } } while (0)

// Host driver for one batch of scans.  d_pts / d_labels: device arrays of the whole batch.
int extract_instances(sgtd_handle *h, const float4 *d_pts, const uint32_t *d_labels, const std::vector<int64_t> &off,
                      int32_t *d_point_instance, std::vector<sgtd_node> &nodes_out, std::vector<int64_t> &node_off,
                      std::vector<int32_t> &n_instances) {
  cudaStream_t st = h->stream;
  const int nscans = (int)off.size() - 1;
  const int64_t total_pts = off[nscans] - off[0];
  int rc = SGTD_OK;
  // option s1_trace: wall-clock checkpoints of the host driver on stderr (experiments)
  const auto t_begin = std::chrono::steady_clock::now();
  auto trace = [&](const char *what) {
    if (!h->opt.s1_trace) return;
    cudaStreamSynchronize(st);
    fprintf(stderr, "[s1] %-28s %8.3f ms\n", what,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
  };
  node_off.assign(nscans + 1, 0);
  n_instances.assign(nscans, 0);
  nodes_out.clear();
  if (nscans == 0) return SGTD_OK;
  // device-side temporaries (freed at the end; stage 1 is not on the per-query path)
  // device-side temporaries live in a per-handle pool: steady-state calls do no cudaMalloc/cudaFree
  S1Pool &sp = s1_pool(h);
  DevBuf<int64_t> &d_off = sp.d_off; DevBuf<uint32_t> &d_cnt = sp.d_cnt; DevBuf<Task> &d_tasks = sp.d_tasks;
  DevBuf<TaskState> &d_ts = sp.d_ts;
  DevBuf<int> &d_pp = sp.d_pp, &d_tab = sp.d_tab, &d_lab = sp.d_lab, &d_pool = sp.d_pool, &d_map = sp.d_map;
  DevBuf<double> &d_polar = sp.d_polar, &d_bounds = sp.d_bounds; DevBuf<unsigned long long> &d_cur = sp.d_cur;
  DevBuf<InstRec> &d_inst = sp.d_inst; DevBuf<sgtd_node> &d_nodes = sp.d_nodes;
  DevBuf<int> &d_map2 = sp.d_map2; DevBuf<uint32_t> &d_sort = sp.d_sort; DevBuf<int64_t> &d_seg = sp.d_seg;
  DevBuf<unsigned char> &d_cub = sp.d_cub;
  std::vector<uint32_t> hc((size_t)nscans * kMaxClass * 2 + 1);
  std::vector<Task> tasks;
  std::vector<TaskState> ts;
  std::vector<int> pool;
  std::vector<InstRec> inst;
  int64_t n_idx = 0, n_tab = 0, n_lab = 0, n_bnd = 0;
  S1Buffers B{};
  unsigned long long cursor = 0;
  {
    S1_CUDA(d_off.reserve(nscans + 1, st, false));
    S1_CUDA(cudaMemcpyAsync(d_off.p, off.data(), (nscans + 1) * 8, cudaMemcpyHostToDevice, st));
    S1_CUDA(d_cnt.reserve(hc.size(), st, false));
    S1_CUDA(cudaMemsetAsync(d_cnt.p, 0, hc.size() * 4, st));
    uint32_t *counts = d_cnt.p, *nonzero = d_cnt.p + (size_t)nscans * kMaxClass, *bad = d_cnt.p + (size_t)nscans * kMaxClass * 2;
    k_s1_hist<<<dim3(nscans, 8), 256, 0, st>>>(d_labels - off[0], d_off.p, counts, nonzero, bad);
    SGTD_LAUNCHED(h);
    S1_CUDA(cudaMemcpyAsync(hc.data(), d_cnt.p, hc.size() * 4, cudaMemcpyDeviceToHost, st));
    S1_CUDA(cudaStreamSynchronize(st));
    if (hc.back()) { rc = sgtd_fail(h, SGTD_E_INVALID, "semantic label >= 32", __FILE__, __LINE__); goto done; }
  }
  trace("histogram + sync");
  // ---- plan the (scan, class) tasks, classes ascending (gen_labels :99-227) ----
  for (int s = 0; s < nscans; ++s)
    for (int c = 0; c < kMaxClass; ++c) {
      const uint32_t cnt = hc[(size_t)s * kMaxClass + c];
      if (!cnt) continue;
      Task t{};
      t.scan = s; t.cls = c; t.pt0 = off[s] - off[0]; t.npts_scan = (int)(off[s + 1] - off[s]); t.npts = (int)cnt;
      // class tables: gen_labels (get_json.cpp:120-209) or, option s1_variant = 1, local_map_creation
      // (local_map.cpp:372-440: class 19 is not skipped, minSeg 400 for the large-surface classes)
      const bool submap = h->opt.s1_variant == 1;
      if (c == 9 || c == 10) t.policy = P_WHOLE;
      else if (c == 0 || c == 1 || c == 2 || c == 3 || c == 6 || c == 7 || c == 8 || c == 14 || (c == 19 && !submap)) continue;
      else if (hc[(size_t)(nscans + s) * kMaxClass + c]) t.policy = P_GTINST;
      else {
        t.policy = P_DCVC;
        t.minSeg = (c == 17 || c == 18 || c == 15) ? 5 : 300;
        if (submap && (c == 10 || c == 11 || c == 12 || c == 14 || c == 16)) t.minSeg = 400;
      }
      t.idx_off = n_idx; n_idx += cnt;
      t.lab_off = n_lab; n_lab += (t.policy == P_GTINST) ? 65536 : (t.policy == P_DCVC ? (int64_t)cnt + 1 : 1);
      if (t.policy == P_DCVC) {
        int sz = 64; while (sz < 2 * (int)cnt) sz <<= 1;
        t.tab_size = sz; t.tab_off = n_tab; n_tab += sz;
        t.bnd_off = n_bnd; n_bnd += 1024;
      }
      tasks.push_back(t);
    }
  trace("task plan (host)");
  if (!tasks.empty()) {
    const int nt = (int)tasks.size();
    ts.resize(nt);
    S1_CUDA(d_tasks.reserve(nt, st, false)); S1_CUDA(d_ts.reserve(nt, st, false));
    S1_CUDA(cudaMemcpyAsync(d_tasks.p, tasks.data(), nt * sizeof(Task), cudaMemcpyHostToDevice, st));
    S1_CUDA(cudaMemsetAsync(d_ts.p, 0, nt * sizeof(TaskState), st));
    S1_CUDA(d_pp.reserve((size_t)std::max<int64_t>(n_idx, 1) * (4 + 4 + 2), st, false));  // cls_idx, slot, pt_label, final | events (int4) | evc (int2)
    S1_CUDA(d_polar.reserve((size_t)std::max<int64_t>(n_idx, 1) * 3, st, false));
    S1_CUDA(d_tab.reserve((size_t)std::max<int64_t>(n_tab, 1) * 7, st, false));  // key,min1,min2,coord,kind,label,first event
    if (h->opt.s1_replay == 0) S1_CUDA(sp.d_rows.reserve((size_t)std::max<int64_t>(n_idx, 1) * 32, st, false));  // neighbour rows, one per event slot
    S1_CUDA(d_lab.reserve((size_t)std::max<int64_t>(n_lab, 1) * 3, st, false));  // parent,count,first
    S1_CUDA(d_bounds.reserve((size_t)std::max<int64_t>(n_bnd, 1), st, false));
    S1_CUDA(d_pool.reserve((size_t)std::max<int64_t>(n_lab, 1) * 3, st, false));
    S1_CUDA(d_cur.reserve(1, st, false));
    S1_CUDA(d_map.reserve((size_t)std::max<int64_t>(n_lab, 1), st, false));
    S1_CUDA(d_map2.reserve((size_t)std::max<int64_t>(n_lab, 1), st, false));  // pool entry of each label
    S1_CUDA(d_sort.reserve((size_t)std::max<int64_t>(n_idx, 1) * 4, st, false));  // key / value double buffers
    S1_CUDA(sp.d_cent.reserve((size_t)std::max<int64_t>(n_lab, 1), st, false));  // centroid per pool entry (sparsely touched)
    B.pts = d_pts; B.labels = d_labels; B.tasks = d_tasks.p; B.ts = d_ts.p;
    B.cls_idx = d_pp.p; B.slot = d_pp.p + n_idx; B.pt_label = d_pp.p + 2 * n_idx; B.final_label = d_pp.p + 3 * n_idx;
    B.events = reinterpret_cast<int4 *>(d_pp.p + 4 * n_idx);  // 4*n_idx ints = 16-byte aligned
    B.evc = reinterpret_cast<int2 *>(d_pp.p + 8 * n_idx);
    B.polar = d_polar.p;
    B.t_key = d_tab.p; B.t_min1 = d_tab.p + n_tab; B.t_min2 = d_tab.p + 2 * n_tab; B.t_coord = d_tab.p + 3 * n_tab;
    B.t_kind = d_tab.p + 4 * n_tab; B.t_label = d_tab.p + 5 * n_tab;
    B.parent = d_lab.p; B.count = d_lab.p + n_lab; B.first = d_lab.p + 2 * n_lab;
    B.bounds = d_bounds.p; B.pool = d_pool.p; B.pool_cursor = d_cur.p; B.pool_of_label = d_map2.p;
    // initial values: keys empty, min1/min2/first = INT_MAX, kind/label/count = 0/-1/0, maps -1
    k_fill_i32<<<1024, 256, 0, st>>>(B.t_key, 3 * n_tab, kEmptyVoxel);  // key, min1, min2
    k_fill_i32<<<1024, 256, 0, st>>>(B.t_kind, n_tab, K_NONE);
    k_fill_i32<<<1024, 256, 0, st>>>(B.t_label, n_tab, -1);
    k_fill_i32<<<1024, 256, 0, st>>>(B.pt_label, n_idx, -1);
    k_fill_i32<<<1024, 256, 0, st>>>(B.parent, n_lab, 0);
    k_fill_i32<<<1024, 256, 0, st>>>(B.count, n_lab, 0);
    k_fill_i32<<<1024, 256, 0, st>>>(B.first, n_lab, kEmptyVoxel);
    k_fill_i32<<<1024, 256, 0, st>>>(d_map.p, n_lab, -1);
    S1_CUDA(cudaMemsetAsync(d_cur.p, 0, 8, st));
    trace("alloc + fills");
    k_s1_gather<<<nt, kS1Threads, 0, st>>>(B);
    k_dcvc_prepare<kS1Threads><<<nt, kS1Threads, 0, st>>>(B, 0.35, 0.0004, 1.2, 1.2, d_tab.p + 6 * n_tab);
    k_dcvc_prepare<1024><<<nt, 1024, 0, st>>>(B, 0.35, 0.0004, 1.2, 1.2, d_tab.p + 6 * n_tab);  // get_json.cpp:205-208
    trace("gather + prepare");
    {
      // table classes of the replay: every class is launched over all tasks, a CTA leaves at once unless its
      // task's voxel count (known on the device only) falls in the class; the last launch takes the rest
      // with the table in global memory.  The classes run concurrently on the pool's side streams.
      static const int kSlots[4] = {2048, 8192, 16384, 32768};
      S1_CUDA(cudaFuncSetAttribute(k_dcvc_replay<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSlots[3] * 6));
      if (!sp.have_streams) {
        for (auto &x : sp.side) S1_CUDA(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
        for (auto &x : sp.ev) S1_CUDA(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
        sp.have_streams = true;
      }
      S1_CUDA(cudaEventRecord(sp.ev[0], st));
      int lo = -1;
      const bool force_global = h->opt.s1_table == 1;  // tests: every task through the global-memory form
      // default: the component-parallel form for every task with at most kCcMaxEv events (three shared-memory
      // classes by event count, on their own streams); option s1_replay = 1: sequential forms only
      const bool use_cc = h->opt.s1_replay == 0 && !force_global;
      static const int kCaps[3] = {2048, 8192, kCcMaxEv};
      if (use_cc) {
        S1_CUDA(cudaFuncSetAttribute(k_dcvc_replay_cc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cc_smem(kCaps[2])));
        k_dcvc_rows<<<dim3(nt, 16), kS1Threads, 0, st>>>(B, d_tab.p + 6 * n_tab, sp.d_rows.p);
        S1_CUDA(cudaEventRecord(sp.ev[0], st));
        int nlo = -1;
        for (int c = 0; c < 3; ++c) {
          S1_CUDA(cudaStreamWaitEvent(sp.side[c], sp.ev[0], 0));
          k_dcvc_replay_cc<<<nt, kCcThreads, cc_smem(kCaps[c]), sp.side[c]>>>(B, sp.d_rows.p, kCaps[c], nlo);
          nlo = kCaps[c];
        }
        h->launches += 4;
      }
      const int cc_max = use_cc ? kCcMaxEv : -1;
      for (int c = 0; c < 4 && !force_global; ++c) {
        const int hi = kSlots[c] * 3 / 10;  // load factor <= 0.3
        cudaStream_t sc = sp.side[c];
        S1_CUDA(cudaStreamWaitEvent(sc, sp.ev[0], 0));
        k_dcvc_replay<true><<<nt, kRpThreads, (size_t)kSlots[c] * 6, sc>>>(B, kSlots[c], lo, hi, h->opt.s1_rows, cc_max);
        S1_CUDA(cudaEventRecord(sp.ev[1 + c], sc));
        lo = hi;
      }
      // the global form (tables too large for shared memory; rare) looks its rows up itself
      k_dcvc_replay<false><<<nt, kRpThreads, 0, st>>>(B, 0, lo, 0x7fffffff, 0, cc_max);
      if (!force_global)
        for (int c = 0; c < 4; ++c) S1_CUDA(cudaStreamWaitEvent(st, sp.ev[1 + c], 0));
    }
    trace("replay");
    k_s1_finish<kS1Threads><<<nt, kS1Threads, 0, st>>>(B);
    k_s1_finish<1024><<<nt, 1024, 0, st>>>(B);
    h->launches += 16;
    S1_CUDA(cudaGetLastError());
    // The label pool goes to the host on a side stream; meanwhile the main stream lists every would-be node
    // instance's points (stable sort by pool index) and sums their centroids -- neither depends on the
    // instance ORDER the host is about to work out.
    S1_CUDA(cudaEventRecord(sp.ev[0], st));
    cudaStream_t sd = sp.side[0];
    S1_CUDA(cudaStreamWaitEvent(sd, sp.ev[0], 0));
    S1_CUDA(cudaMemcpyAsync(ts.data(), d_ts.p, nt * sizeof(TaskState), cudaMemcpyDeviceToHost, sd));
    S1_CUDA(cudaMemcpyAsync(&cursor, d_cur.p, 8, cudaMemcpyDeviceToHost, sd));
    uint32_t *k0 = d_sort.p, *k1 = d_sort.p + n_idx, *v0 = d_sort.p + 2 * n_idx, *v1 = d_sort.p + 3 * n_idx;
    {
      k_s1_lkey<<<nt, kS1Threads, 0, st>>>(B, k0, v0);
      int bits = 1;  // pool indices are below the number of label slots
      while ((1ll << bits) <= n_lab) ++bits;
      size_t cb = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, cb, k0, k1, v0, v1, (int)n_idx, 0, bits, st);
      S1_CUDA(d_cub.reserve(cb, st, false));
      // (all-ones keys -- points outside node instances -- are the largest value in the low `bits` bits too:
      // they end up behind every pool index)
      S1_CUDA(cub::DeviceRadixSort::SortPairs(d_cub.p, cb, k0, k1, v0, v1, (int)n_idx, 0, std::min(32, (bits + 7) / 8 * 8), st));
      k_s1_lcentroid<<<h->sm_count * 32, 128, 0, st>>>(B, k1, v1, n_idx, sp.d_cent.p);
      h->launches += 3;
      S1_CUDA(cudaGetLastError());
    }
    S1_CUDA(cudaStreamSynchronize(sd));
    pool.resize((size_t)cursor * 3 + 3);
    if (cursor) S1_CUDA(cudaMemcpyAsync(pool.data(), d_pool.p, (size_t)cursor * 12, cudaMemcpyDeviceToHost, sd));
    S1_CUDA(cudaStreamSynchronize(sd));
    trace("finish + D2H labels");
    if (h->opt.s1_trace) {
      std::vector<int> ord(nt);
      for (int i = 0; i < nt; ++i) ord[i] = i;
      std::sort(ord.begin(), ord.end(), [&](int a, int b) { return ts[a].dbg_cyc[2] > ts[b].dbg_cyc[2]; });
      for (int j = 0; j < (h->opt.s1_trace >= 2 ? nt : std::min(nt, 6)); ++j) {
        const TaskState &x = ts[ord[j]];
        fprintf(stderr, "[s1] task %4d cls %2d npts %6d nev %6d nvox %5d labels %5d active %5d windows %5d | kcyc build %lld transl %lld replay %lld wb %lld\n",
                ord[j], tasks[ord[j]].cls, tasks[ord[j]].npts, x.nevents, x.nvox, x.labelCount, x.dbg_active, x.dbg_windows,
                x.dbg_cyc[0] / 1000, x.dbg_cyc[1] / 1000, x.dbg_cyc[2] / 1000, x.dbg_cyc[3] / 1000);
      }
    }
    // ---- instance ids: per scan, tasks in class order; cluster order per policy ----
    // The cluster order of a DCVC task is the iteration order of the reference's
    // std::unordered_map<int, vector<int>> label2segIndex (labelAnalysis, :394-418), keyed by label and filled
    // in ascending point order.  That order only depends on the key sequence, so the real container is
    // replayed with the labels in first-appearance order (mapped type irrelevant).  Tasks are independent:
    // they are ordered on a few host threads.
    struct L { int label, count, first, pidx; };
    std::vector<std::vector<L>> orders((size_t)nt);
    auto order_task = [&](int ti) {
      const Task &t = tasks[ti];
      std::vector<L> ls((size_t)ts[ti].ndistinct);
      for (int j = 0; j < ts[ti].ndistinct; ++j) {
        const int *p = &pool[(size_t)(ts[ti].pool_off + j) * 3];
        ls[j] = L{p[0], p[1], p[2], (int)(ts[ti].pool_off + j)};
      }
      std::vector<L> &order = orders[(size_t)ti];
      if (t.policy == P_DCVC) {
        std::sort(ls.begin(), ls.end(), [](const L &a, const L &b) { return a.first < b.first; });
        std::unordered_map<int, int> label2segIndex;  // value: position in ls
        for (size_t j = 0; j < ls.size(); ++j) label2segIndex[ls[j].label] = (int)j;
        for (auto &it : label2segIndex)
          if (ls[(size_t)it.second].count >= t.minSeg) order.push_back(ls[(size_t)it.second]);
      } else if (t.policy == P_GTINST) {
        std::sort(ls.begin(), ls.end(), [](const L &a, const L &b) { return a.label < b.label; });
        for (const L &l : ls) if (l.count > 20) order.push_back(l);  // get_json.cpp:146
      } else {
        order = ls;
      }
    };
    {
      const int nthr = std::max(1, std::min(8, (int)std::thread::hardware_concurrency()));
      if (nthr == 1 || nt < 64) {
        for (int ti = 0; ti < nt; ++ti) order_task(ti);
      } else {
        std::vector<std::thread> pool_thr;
        std::atomic<int> next{0};
        for (int w = 0; w < nthr; ++w)
          pool_thr.emplace_back([&] { for (int ti = next.fetch_add(8); ti < nt; ti = next.fetch_add(8)) for (int j = ti; j < std::min(nt, ti + 8); ++j) order_task(j); });
        for (auto &th : pool_thr) th.join();
      }
    }
    int cur_scan = -1, inst_id = 0;
    int64_t node_cursor = 0;
    for (int ti = 0; ti < nt; ++ti) {
      const Task &t = tasks[ti];
      if (t.scan != cur_scan) {
        if (cur_scan >= 0) n_instances[cur_scan] = inst_id;
        for (int s = cur_scan + 1; s <= t.scan; ++s) node_off[s] = node_cursor;
        cur_scan = t.scan; inst_id = 0;
      }
      const std::vector<L> &order = orders[(size_t)ti];
      const int mapped = node_map(t.cls);
      for (const L &l : order) {
        InstRec ir{};
        ir.task = ti; ir.label = l.label; ir.inst_id = inst_id++;
        ir.node_slot = -1; ir.node_label = 0;
        ir.pool_idx = l.pidx;
        if (mapped >= 3 && mapped <= 12) { ir.node_slot = (int)node_cursor++; ir.node_label = (uint32_t)mapped; }
        inst.push_back(ir);
      }
    }
    if (cur_scan >= 0) n_instances[cur_scan] = inst_id;
    for (int s = cur_scan + 1; s <= nscans; ++s) node_off[s] = node_cursor;
    nodes_out.resize((size_t)node_cursor);
    trace("instance order (host)");
    if (!inst.empty()) {
      const int ni = (int)inst.size();
      S1_CUDA(d_inst.reserve(ni, st, false));
      S1_CUDA(cudaMemcpyAsync(d_inst.p, inst.data(), ni * sizeof(InstRec), cudaMemcpyHostToDevice, st));
      S1_CUDA(d_nodes.reserve((size_t)std::max<int64_t>(node_cursor, 1), st, false));
      k_s1_scatter_map<<<(ni + 255) / 256, 256, 0, st>>>(B, d_inst.p, ni, d_map.p, sp.d_cent.p, d_nodes.p);
      if (d_point_instance) k_s1_assign<<<nt, kS1Threads, 0, st>>>(B, d_map.p, d_point_instance);
      h->launches += 2;
      S1_CUDA(cudaGetLastError());
      if (node_cursor) S1_CUDA(cudaMemcpyAsync(nodes_out.data(), d_nodes.p, (size_t)node_cursor * sizeof(sgtd_node), cudaMemcpyDeviceToHost, st));
    }
    S1_CUDA(cudaStreamSynchronize(st));
    trace("assign + centroid + D2H");
  }
  (void)total_pts;
done:
  cudaStreamSynchronize(st);
  return rc;
}

}  // namespace sgtd

using namespace sgtd;

static bool s1_is_device_ptr(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

extern "C" int sgtd_extract_instances_batch(sgtd_handle *h, const float *points, const uint32_t *labels,
                                            const int64_t *scan_offsets, int32_t nscans, int32_t *point_instance,
                                            sgtd_node *nodes, int64_t cap_nodes, int64_t *node_offsets,
                                            int32_t *n_instances) {
  if (!h || nscans < 0 || (nscans > 0 && (!points || !labels || !scan_offsets)) || !node_offsets)
    return sgtd_fail(h, SGTD_E_INVALID, "bad argument", __FILE__, __LINE__);
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(h->device);
  cudaStream_t st = h->stream;
  std::vector<int64_t> off(scan_offsets, scan_offsets + nscans + 1);
  const int64_t total = nscans ? off[nscans] - off[0] : 0;
  S1Pool &spool = s1_pool(h);
  DevBuf<float4> &dp = spool.in_pts; DevBuf<uint32_t> &dl = spool.in_lab; DevBuf<int32_t> &dpi = spool.out_pi;
  const float4 *d_pts = reinterpret_cast<const float4 *>(points) + (nscans ? off[0] : 0);
  const uint32_t *d_lab = labels + (nscans ? off[0] : 0);
  int rc = SGTD_OK;
  std::vector<sgtd_node> nodes_out; std::vector<int64_t> noff; std::vector<int32_t> ninst;
  cudaError_t e = cudaSuccess;
  if (total > 0 && !s1_is_device_ptr(points)) {
    if ((e = dp.reserve((size_t)total, st, false)) == cudaSuccess)
      e = cudaMemcpyAsync(dp.p, d_pts, (size_t)total * 16, cudaMemcpyHostToDevice, st);
    d_pts = dp.p;
  }
  if (e == cudaSuccess && total > 0 && !s1_is_device_ptr(labels)) {
    if ((e = dl.reserve((size_t)total, st, false)) == cudaSuccess)
      e = cudaMemcpyAsync(dl.p, d_lab, (size_t)total * 4, cudaMemcpyHostToDevice, st);
    d_lab = dl.p;
  }
  int32_t *d_pi = nullptr;
  const bool pi_dev = point_instance && s1_is_device_ptr(point_instance);
  if (e == cudaSuccess && point_instance && total > 0) {
    if (pi_dev) d_pi = point_instance + off[0];
    else { e = dpi.reserve((size_t)total, st, false); d_pi = dpi.p; }
    if (e == cudaSuccess) e = cudaMemsetAsync(d_pi, 0xFF, (size_t)total * 4, st);
  }
  if (e != cudaSuccess) rc = sgtd_fail(h, SGTD_E_CUDA, "stage-1 staging", __FILE__, __LINE__, e);
  if (rc == SGTD_OK) rc = extract_instances(h, d_pts, d_lab, off, d_pi, nodes_out, noff, ninst);
  if (rc == SGTD_OK) {
    if ((int64_t)nodes_out.size() > cap_nodes) rc = sgtd_fail(h, SGTD_E_CAPACITY, "node buffer too small", __FILE__, __LINE__);
    else {
      if (!nodes_out.empty()) {
        if (s1_is_device_ptr(nodes)) cudaMemcpy(nodes, nodes_out.data(), nodes_out.size() * sizeof(sgtd_node), cudaMemcpyHostToDevice);
        else memcpy(nodes, nodes_out.data(), nodes_out.size() * sizeof(sgtd_node));
      }
      memcpy(node_offsets, noff.data(), (nscans + 1) * 8);
      if (n_instances && nscans) memcpy(n_instances, ninst.data(), nscans * 4);
      if (point_instance && !pi_dev && total > 0) {
        e = cudaMemcpy(point_instance + off[0], d_pi, (size_t)total * 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = sgtd_fail(h, SGTD_E_CUDA, "copy point_instance", __FILE__, __LINE__, e);
      }
    }
  }
  if (prev >= 0 && prev != h->device) cudaSetDevice(prev);
  return rc;
}

extern "C" int sgtd_extract_instances(sgtd_handle *h, const float *points, const uint32_t *labels, int64_t n,
                                      int32_t *point_instance, sgtd_node *nodes, int32_t cap_nodes, int32_t *n_nodes,
                                      int32_t *n_instances) {
  const int64_t off[2] = {0, n};
  int64_t noff[2] = {0, 0};
  int32_t ni = 0;
  int rc = sgtd_extract_instances_batch(h, points, labels, off, 1, point_instance, nodes, cap_nodes, noff, &ni);
  if (rc) return rc;
  if (n_nodes) *n_nodes = (int32_t)noff[1];
  if (n_instances) *n_instances = ni;
  return SGTD_OK;
}

// ---- submap aggregation: the point-gathering part of local_map_creation -----------------
// R/src/local_map.cpp:213-328, literally: the scan's own points, then one transformed copy per other scan of
// the submap within `radius` of it -- of the CURRENT scan's points (the reference re-opens current_scan_path
// for every neighbour, :272), the intensity acting as homogeneous coordinate except for the last point, whose
// four components are one (:290).  A pure streaming transform: 16 B read (L2 after the first copy) and 20 B
// written per output point.
namespace sgtd {
struct Mat44f { float m[16]; };
__global__ void k_submap_copies(const float4 *pts, const uint32_t *lab, int64_t n, const Mat44f *T, int ncopies, float4 *out,
                                uint32_t *out_lab) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const float4 v0 = pts[p];
  const uint32_t l = lab[p];
  out[p] = v0; out_lab[p] = l;
  float4 v = v0;
  if (p == n - 1) v = make_float4(1.f, 1.f, 1.f, 1.f);
  for (int c = 0; c < ncopies; ++c) {
    const float *m = T[c].m;
    float4 o;
    o.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], v.x), __fmul_rn(m[1], v.y)), __fmul_rn(m[2], v.z)), __fmul_rn(m[3], v.w));
    o.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[4], v.x), __fmul_rn(m[5], v.y)), __fmul_rn(m[6], v.z)), __fmul_rn(m[7], v.w));
    o.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[8], v.x), __fmul_rn(m[9], v.y)), __fmul_rn(m[10], v.z)), __fmul_rn(m[11], v.w));
    o.w = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[12], v.x), __fmul_rn(m[13], v.y)), __fmul_rn(m[14], v.z)), __fmul_rn(m[15], v.w));
    out[(int64_t)(c + 1) * n + p] = o;
    out_lab[(int64_t)(c + 1) * n + p] = l;
  }
}
static void mul44f(const float *a, const float *b, float *c) {  // Matrix4f product, k = 0..3 in order
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float v = a[i * 4] * b[j];
      v += a[i * 4 + 1] * b[4 + j];
      v += a[i * 4 + 2] * b[8 + j];
      v += a[i * 4 + 3] * b[12 + j];
      c[i * 4 + j] = v;
    }
}
static void inv44f(const float *m, float *o) {  // last row (0,0,0,1): cofactors of the 3x3 block
  const float c00 = m[5] * m[10] - m[6] * m[9], c01 = m[6] * m[8] - m[4] * m[10], c02 = m[4] * m[9] - m[5] * m[8];
  const float id = 1.0f / (m[0] * c00 + m[1] * c01 + m[2] * c02);
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
}  // namespace sgtd

extern "C" int sgtd_submap_aggregate(sgtd_handle *h, const float *points, const uint32_t *labels, int64_t n,
                                     const float *poses12, int32_t nscans, int32_t j, const float *base2ouster16,
                                     float radius, float *out_points, uint32_t *out_labels, int64_t cap, int64_t *n_out,
                                     int32_t *n_used) {
  if (!h || !points || !labels || !poses12 || !n_out || n < 0 || nscans < 1 || j < 0 || j >= nscans)
    return sgtd_fail(h, SGTD_E_INVALID, "bad argument", __FILE__, __LINE__);
  float Tj[16], Tji[16], I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  auto pose44 = [&](int s, float *m) { for (int k = 0; k < 12; ++k) m[k] = poses12[(size_t)s * 12 + k]; m[12] = m[13] = m[14] = 0.0f; m[15] = 1.0f; };
  pose44(j, Tj);
  inv44f(Tj, Tji);
  const float *B = base2ouster16 ? base2ouster16 : I;
  std::vector<Mat44f> Ts;
  for (int i = 0; i < nscans; ++i) {
    if (i == j) continue;
    const float dx = Tj[3] - poses12[(size_t)i * 12 + 3], dy = Tj[7] - poses12[(size_t)i * 12 + 7], dz = Tj[11] - poses12[(size_t)i * 12 + 11];
    if (sqrtf(dx * dx + (dy * dy + dz * dz)) > radius) continue;  // (t1 - t2).norm() > 15 (:268)
    float Ti[16], A[16];
    Mat44f T;
    pose44(i, Ti);
    mul44f(Tji, Ti, A);  // T_j.inverse() * T_i * BASE2OUSTER (:292)
    mul44f(A, B, T.m);
    Ts.push_back(T);
  }
  const int64_t total = n * (int64_t)(Ts.size() + 1);
  *n_out = total;
  if (n_used) *n_used = (int32_t)Ts.size() + 1;
  if (!out_points || !out_labels || total > cap) return total > cap ? SGTD_E_CAPACITY : SGTD_OK;
  if (total == 0) return SGTD_OK;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(h->device);
  cudaStream_t st = h->stream;
  S1Pool &sp = s1_pool(h);
  int rc = SGTD_OK;
  cudaError_t e = cudaSuccess;
  const bool in_dev = s1_is_device_ptr(points), out_dev = s1_is_device_ptr(out_points);
  const float4 *d_pts = reinterpret_cast<const float4 *>(points);
  const uint32_t *d_lab = labels;
  DevBuf<float4> d_out; DevBuf<uint32_t> d_outl; DevBuf<Mat44f> d_T;
  do {
    if (!in_dev) {
      if ((e = sp.in_pts.reserve((size_t)n, st, false)) != cudaSuccess) break;
      if ((e = sp.in_lab.reserve((size_t)n, st, false)) != cudaSuccess) break;
      if ((e = cudaMemcpyAsync(sp.in_pts.p, points, (size_t)n * 16, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
      if ((e = cudaMemcpyAsync(sp.in_lab.p, labels, (size_t)n * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
      d_pts = sp.in_pts.p; d_lab = sp.in_lab.p;
    }
    float4 *o_pts = reinterpret_cast<float4 *>(out_points);
    uint32_t *o_lab = out_labels;
    if (!out_dev) {
      if ((e = d_out.reserve((size_t)total, st, false)) != cudaSuccess) break;
      if ((e = d_outl.reserve((size_t)total, st, false)) != cudaSuccess) break;
      o_pts = d_out.p; o_lab = d_outl.p;
    }
    if ((e = d_T.reserve(std::max<size_t>(Ts.size(), 1), st, false)) != cudaSuccess) break;
    if (!Ts.empty() && (e = cudaMemcpyAsync(d_T.p, Ts.data(), Ts.size() * sizeof(Mat44f), cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    k_submap_copies<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_pts, d_lab, n, d_T.p, (int)Ts.size(), o_pts, o_lab);
    SGTD_LAUNCHED(h);
    if ((e = cudaGetLastError()) != cudaSuccess) break;
    if (!out_dev) {
      if ((e = cudaMemcpyAsync(out_points, o_pts, (size_t)total * 16, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
      if ((e = cudaMemcpyAsync(out_labels, o_lab, (size_t)total * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
    }
    e = cudaStreamSynchronize(st);
  } while (false);
  if (e != cudaSuccess) rc = sgtd_fail(h, SGTD_E_CUDA, "sgtd_submap_aggregate", __FILE__, __LINE__, e);
  d_out.release(); d_outl.release(); d_T.release();
  if (prev >= 0 && prev != h->device) cudaSetDevice(prev);
  return rc;
}
