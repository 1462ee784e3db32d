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
// atomicCAS/atomicMin, event list by ordered compaction), then one warp per
// (scan, class) task replays the event list: lanes 0..26 look the 27 neighbour voxels
// up in parallel, lane 0 applies the walk.  Many tasks run concurrently (one scan has
// ~8 class tasks; a batch of scans fills the machine).
// The cluster -> instance-id order is the iteration order of the reference's
// std::unordered_map<int, vector<int>> (cluster_manager.hpp:395-418); it is reproduced
// on the host with the real container fed the labels in first-appearance order.
// Centroids are sequential float32 sums in ascending point order (get_json.cpp:266-274).
#include <algorithm>
#include <cmath>
#include <unordered_map>

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
  int64_t pool_off;
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
  int *nbr;          // 27 neighbour slots per event
  int *pt_label;     // labels of points of invisible voxels
  int *final_label;  // final label per point
  int *parent, *count, *first;  // per label value
  int *t_key, *t_min1, *t_min2, *t_coord, *t_kind, *t_label;  // voxel table
  double *bounds;
  int *pool;         // (label, count, first) triples
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

__device__ __forceinline__ int block_excl_scan(int v, int *s_warp, int &total) {
  // exclusive scan over the 256 threads of the block (in thread order)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  __syncthreads();
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int base = 0; total = 0;
#pragma unroll
  for (int w = 0; w < kS1Threads / 32; ++w) { if (w < warp) base += s_warp[w]; total += s_warp[w]; }
  return base + incl - v;
}

// ---- K2: ordered list of the scan's points that carry the task's class ----------------
__global__ void __launch_bounds__(kS1Threads) k_s1_gather(S1Buffers B) {
  __shared__ int s_warp[kS1Threads / 32];
  const Task t = B.tasks[blockIdx.x];
  int base = 0;
  for (int i0 = 0; i0 < t.npts_scan; i0 += kS1Threads) {
    const int i = i0 + threadIdx.x;
    const int hit = (i < t.npts_scan) && ((B.labels[t.pt0 + i] & 0xFFFFu) == (uint32_t)t.cls);
    int total;
    const int pos = block_excl_scan(hit, s_warp, total);
    if (hit) B.cls_idx[t.idx_off + base + pos] = i;
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
__global__ void __launch_bounds__(kS1Threads) k_dcvc_prepare(S1Buffers B, double startR, double deltaR, double deltaP,
                                                              double deltaA) {
  __shared__ int s_warp[kS1Threads / 32];
  __shared__ double s_red[4][kS1Threads / 32];
  __shared__ double s_mm[4];
  __shared__ int s_grid[3];
  const Task t = B.tasks[blockIdx.x];
  if (t.policy != P_DCVC) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double kPi = 3.14159265358979323846;  // M_PI
  double *polar = B.polar + 3 * t.idx_off;
  // 1. convert2polar (cluster_manager.hpp:172-206); running min/max start at 0,0,5,5 (:482-485)
  double mnP = 0.0, mxP = 0.0, mnR = 5.0, mxR = 5.0;
  for (int r = tid; r < t.npts; r += kS1Threads) {
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
    for (int w = 1; w < kS1Threads / 32; ++w) {
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
  for (int r = tid; r < t.npts; r += kS1Threads) {
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
  for (int r = tid; r < t.npts; r += kS1Threads) {
    const int pos = B.slot[t.idx_off + r];
    if (__ldcg(&t_min1[pos]) != r) atomicMin(&t_min2[pos], r);
  }
  __syncthreads();
  // 3. seed events in point order: two lowest points of a visible voxel, every point of an invisible one
  int base = 0;
  for (int r0 = 0; r0 < t.npts; r0 += kS1Threads) {
    const int r = r0 + tid;
    int ev = 0;
    if (r < t.npts) {
      const int pos = B.slot[t.idx_off + r];
      const int pitchIndex = __ldcg(&t_coord[pos]) >> 21;
      const bool visible = pitchIndex <= height;
      ev = !visible || __ldcg(&t_min1[pos]) == r || __ldcg(&t_min2[pos]) == r;
    }
    int total;
    const int p = block_excl_scan(ev, s_warp, total);
    if (ev) {
      const int pos = B.slot[t.idx_off + r];
      const int which = (__ldcg(&t_min1[pos]) == r ? 1 : 0) | (__ldcg(&t_min2[pos]) == r ? 2 : 0);
      B.events[t.idx_off + base + p] = make_int4(r, pos, __ldcg(&t_coord[pos]), which);
    }
    base += total;
  }
  if (tid == 0) B.ts[blockIdx.x].nevents = base;
  __syncthreads();
  // 4. neighbour voxel slots of every event, searchKNN order (:365-385): z (pitch) outer, y (polar),
  //    x (azimuth) inner; -1 = skipped by the guards or no such voxel.  Parallel here, so that the
  //    sequential replay only has to read 27 consecutive ints per event.
  const int nev = base;
  int *nbr = B.nbr + 27 * t.idx_off;
  for (int w = tid; w < nev * 27; w += kS1Threads) {
    const int e = w / 27, k = w - e * 27;
    const int coord = B.events[t.idx_off + e].z;
    const int az = coord & 1023, po = (coord >> 10) & 2047, pi = coord >> 21;
    const int z = pi - 1 + k / 9, y = po - 1 + (k / 3) % 3, x = az - 1 + k % 3;
    int nb = -1;
    if (!(z < 0 || z > height) && !(y < 0 || y > polarNum)) {
      int ax = x;
      if (ax < 0) ax = width - 1;
      if (ax > 300) ax = 300;
      nb = table_lookup(t_key, mask, (ax * (polarNum + 1) + y) + z * (polarNum + 1) * (width + 1));
    }
    nbr[w] = nb;
  }
}

__device__ __forceinline__ int uf_find_ro(const int *parent, int x) {
  while (true) { const int p = __ldcg(parent + x); if (p == x) return x; x = p; }
}

// Union-find over label values: the first kUfSmem labels live in shared memory (every task of the
// test and bench workloads fits), the rest in the task's global array.
constexpr int kUfSmem = 8192;
struct UnionFind {
  int *sm, *gl;
  __device__ __forceinline__ int get(int x) const { return x < kUfSmem ? sm[x] : __ldcg(gl + x); }
  __device__ __forceinline__ void set(int x, int v) { if (x < kUfSmem) sm[x] = v; else gl[x] = v; }
  __device__ __forceinline__ int find(int x) {
    int r = x;
    while (true) { const int p = get(r); if (p == r) break; r = p; }
    while (x != r) { const int p = get(x); set(x, r); x = p; }  // path compression
    return r;
  }
};

// ---- K4: sequential replay of the seed events, one warp per task (DCVC, :272-355) ---------
// Per event the warp reads one prefetched record + 27 precomputed neighbour slots, gathers the 27
// voxel states in parallel, and lane 0 applies the reference's walk.
__global__ void __launch_bounds__(32) k_dcvc_replay(S1Buffers B) {
  __shared__ int s_nb[27], s_st[27], s_lb[27];
  __shared__ int s_parent[kUfSmem];
  const Task t = B.tasks[blockIdx.x];
  if (t.policy != P_DCVC) return;
  const int lane = threadIdx.x;
  const TaskState ts = B.ts[blockIdx.x];
  const int height = ts.height;
  int *t_kind = B.t_kind + t.tab_off, *t_label = B.t_label + t.tab_off;
  UnionFind uf{s_parent, B.parent + t.lab_off};
  int *pt_label = B.pt_label + t.idx_off;
  const int4 *events = B.events + t.idx_off;
  const int *nbr = B.nbr + 27 * t.idx_off;
  int labelCount = 0;
  const int nev = ts.nevents;
  // software pipeline: record and neighbour row of event e+1 are in flight while e is processed
  int4 ev_next = nev > 0 ? events[0] : make_int4(0, 0, 0, 0);
  int nb_next = (nev > 0 && lane < 27) ? nbr[lane] : -1;
  for (int e = 0; e < nev; ++e) {
    const int4 ev = ev_next;
    const int nb_cur = nb_next;
    if (e + 1 < nev) {
      ev_next = events[e + 1];
      if (lane < 27) nb_next = nbr[(e + 1) * 27 + lane];
    }
    const int r = ev.x, v = ev.y;
    const bool vis = (ev.z >> 21) <= height;
    // one round trip per event: the 27 neighbour states (a visible voxel is its own neighbour 13)
    const int st_cur = (lane < 27 && nb_cur >= 0) ? __ldcg(t_kind + nb_cur) : K_NONE;
    const int lb_cur = (lane < 27 && nb_cur >= 0) ? __ldcg(t_label + nb_cur) : -1;
    if (vis) {  // is the point still unlabelled?  (ev.w: bit0 = lowest point of its voxel, bit1 = second lowest)
      const int kd = __shfl_sync(0xffffffffu, st_cur, 13);
      if (kd == K_ALL) continue;
      if (kd == K_NONE && !(ev.w & 1)) continue;
      if (kd == K_HEAD && !(ev.w & 2)) continue;
    }
    if (lane < 27) { s_nb[lane] = nb_cur; s_st[lane] = st_cur; s_lb[lane] = lb_cur; }
    __syncwarp();
    if (lane == 0) {
      int cur = -1;
      bool self_all = false;
      for (int k = 0; k < 27; ++k) {
        const int nb = s_nb[k];
        if (nb < 0) continue;
        const int st = s_st[k];
        if (st == K_NONE) {
          if (cur != -1) { t_kind[nb] = K_ALL; t_label[nb] = cur; if (nb == v) self_all = true; }
        } else {
          const int lab = uf.find(s_lb[k]);
          if (cur == -1) cur = lab;
          else if (cur != lab) { uf.set(cur, lab); cur = lab; }  // relabel sweep cur -> neigh (:323-327)
          if (st == K_HEAD) t_kind[nb] = K_ALL;
          if (nb == v) self_all = true;
        }
      }
      if (cur == -1) {  // new label for the seed and every neighbour (:340-346)
        const int L = ++labelCount;
        uf.set(L, L);
        for (int k = 0; k < 27; ++k) { const int nb = s_nb[k]; if (nb >= 0) { t_kind[nb] = K_ALL; t_label[nb] = L; } }
        if (!vis) pt_label[r] = L;
      } else if (vis) {
        if (!self_all) { t_kind[v] = K_HEAD; t_label[v] = cur; }  // own voxel was passed while cur == -1
      } else {
        pt_label[r] = cur;
      }
    }
    __syncwarp();
  }
  // publish the shared-memory part of the union-find for k_s1_finish
  __syncwarp();
  labelCount = __shfl_sync(0xffffffffu, labelCount, 0);  // only lane 0 counted
  for (int i = 1 + lane; i <= labelCount && i < kUfSmem; i += 32) uf.gl[i] = s_parent[i];
  if (lane == 0) B.ts[blockIdx.x].labelCount = labelCount;
}

// ---- K5: final label per point, per-label size and first point, compact label list -------
__global__ void __launch_bounds__(kS1Threads) k_s1_finish(S1Buffers B) {
  __shared__ int s_warp[kS1Threads / 32];
  __shared__ long long s_pool;
  const Task t = B.tasks[blockIdx.x];
  const int tid = threadIdx.x;
  int *count = B.count + t.lab_off, *first = B.first + t.lab_off;
  int nlab = 0;
  if (t.policy == P_DCVC) {
    const TaskState ts = B.ts[blockIdx.x];
    const int *parent = B.parent + t.lab_off;
    const int *t_coord = B.t_coord + t.tab_off, *t_label = B.t_label + t.tab_off;
    for (int r = tid; r < t.npts; r += kS1Threads) {
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
    for (int r = tid; r < t.npts; r += kS1Threads) {
      const int lab = (int)(B.labels[t.pt0 + B.cls_idx[t.idx_off + r]] >> 16);
      B.final_label[t.idx_off + r] = lab;
      atomicAdd(&count[lab], 1);
      atomicMin(&first[lab], r);
    }
    nlab = 65536;
  } else {
    for (int r = tid; r < t.npts; r += kS1Threads) B.final_label[t.idx_off + r] = 0;
    if (tid == 0) { count[0] = t.npts; first[0] = 0; }
    nlab = 1;
  }
  __syncthreads();
  // distinct labels -> pool (label, count, first); order is fixed on the host
  int mine = 0;
  for (int l = tid; l < nlab; l += kS1Threads) mine += __ldcg(&count[l]) > 0;
  int total;
  int pos = block_excl_scan(mine, s_warp, total);
  if (tid == 0) {
    s_pool = (long long)atomicAdd(B.pool_cursor, (unsigned long long)total);
    B.ts[blockIdx.x].pool_off = s_pool;
    B.ts[blockIdx.x].ndistinct = total;
  }
  __syncthreads();
  for (int l = tid; l < nlab; l += kS1Threads) {
    const int c = __ldcg(&count[l]);
    if (c > 0) {
      int *o = B.pool + 3 * (s_pool + pos);
      o[0] = l; o[1] = c; o[2] = __ldcg(&first[l]);
      ++pos;
    }
  }
}

struct InstRec {  // one per instance, host-planned
  int task, label, inst_id, node_slot;  // node_slot: index into the node output or -1
  uint32_t node_label;
};

// ---- K6: label -> instance id, per point membership ------------------------------------------
__global__ void k_s1_scatter_map(S1Buffers B, const InstRec *inst, int ninst, int *inst_of_label) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ninst) return;
  const Task t = B.tasks[inst[i].task];
  inst_of_label[t.lab_off + inst[i].label] = inst[i].inst_id;
}
__global__ void __launch_bounds__(kS1Threads) k_s1_assign(S1Buffers B, const int *inst_of_label, int32_t *point_instance) {
  const Task t = B.tasks[blockIdx.x];
  for (int r = threadIdx.x; r < t.npts; r += kS1Threads) {
    const int id = inst_of_label[t.lab_off + B.final_label[t.idx_off + r]];
    if (id >= 0) point_instance[t.pt0 + B.cls_idx[t.idx_off + r]] = id;
  }
}

// ---- K7: centroid = sequential float32 sum in ascending point order (get_json.cpp:266-274) ----
// One warp per instance: 32 class points are fetched per trip (coalesced), the lanes that belong to
// the instance are then added in lane order, so the rounding sequence is the reference's.
__global__ void __launch_bounds__(128) k_s1_centroid(S1Buffers B, const InstRec *inst, int ninst, sgtd_node *nodes) {
  const int i = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= ninst) return;
  const InstRec ir = inst[i];
  if (ir.node_slot < 0) return;
  const Task t = B.tasks[ir.task];
  float cx = 0.f, cy = 0.f, cz = 0.f;
  int n = 0;
  for (int r0 = 0; r0 < t.npts; r0 += 32) {
    const int r = r0 + lane;
    const bool mine = r < t.npts && B.final_label[t.idx_off + r] == ir.label;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (mine) p = B.pts[t.pt0 + B.cls_idx[t.idx_off + r]];
    unsigned m = __ballot_sync(0xffffffffu, mine);
    n += __popc(m);
    while (m) {
      const int l = __ffs(m) - 1;
      m &= m - 1;
      cx = __fadd_rn(cx, __shfl_sync(0xffffffffu, p.x, l));
      cy = __fadd_rn(cy, __shfl_sync(0xffffffffu, p.y, l));
      cz = __fadd_rn(cz, __shfl_sync(0xffffffffu, p.z, l));
    }
  }
  if (lane == 0) {
    const float cnt = (float)n;
    sgtd_node nd;
    nd.x = __fdiv_rn(cx, cnt); nd.y = __fdiv_rn(cy, cnt); nd.z = __fdiv_rn(cz, cnt);
    nd.label = ir.node_label;
    nodes[ir.node_slot] = nd;
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
  DevBuf<float4> in_pts; DevBuf<uint32_t> in_lab; DevBuf<int32_t> out_pi;  // staging of host inputs / outputs
  ~S1Pool() {
    d_off.release(); d_cnt.release(); d_tasks.release(); d_ts.release(); d_pp.release(); d_tab.release(); d_lab.release();
    d_pool.release(); d_map.release(); d_polar.release(); d_bounds.release(); d_cur.release(); d_inst.release();
    d_nodes.release(); in_pts.release(); in_lab.release(); out_pi.release();
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
  // ---- plan the (scan, class) tasks, classes ascending (gen_labels :99-227) ----
  for (int s = 0; s < nscans; ++s)
    for (int c = 0; c < kMaxClass; ++c) {
      const uint32_t cnt = hc[(size_t)s * kMaxClass + c];
      if (!cnt) continue;
      Task t{};
      t.scan = s; t.cls = c; t.pt0 = off[s] - off[0]; t.npts_scan = (int)(off[s + 1] - off[s]); t.npts = (int)cnt;
      if (c == 9 || c == 10) t.policy = P_WHOLE;
      else if (c == 0 || c == 1 || c == 2 || c == 3 || c == 6 || c == 7 || c == 8 || c == 14 || c == 19) continue;
      else if (hc[(size_t)(nscans + s) * kMaxClass + c]) t.policy = P_GTINST;
      else { t.policy = P_DCVC; t.minSeg = (c == 17 || c == 18 || c == 15) ? 5 : 300; }
      t.idx_off = n_idx; n_idx += cnt;
      t.lab_off = n_lab; n_lab += (t.policy == P_GTINST) ? 65536 : (t.policy == P_DCVC ? (int64_t)cnt + 1 : 1);
      if (t.policy == P_DCVC) {
        int sz = 64; while (sz < 2 * (int)cnt) sz <<= 1;
        t.tab_size = sz; t.tab_off = n_tab; n_tab += sz;
        t.bnd_off = n_bnd; n_bnd += 1024;
      }
      tasks.push_back(t);
    }
  if (!tasks.empty()) {
    const int nt = (int)tasks.size();
    ts.resize(nt);
    S1_CUDA(d_tasks.reserve(nt, st, false)); S1_CUDA(d_ts.reserve(nt, st, false));
    S1_CUDA(cudaMemcpyAsync(d_tasks.p, tasks.data(), nt * sizeof(Task), cudaMemcpyHostToDevice, st));
    S1_CUDA(cudaMemsetAsync(d_ts.p, 0, nt * sizeof(TaskState), st));
    S1_CUDA(d_pp.reserve((size_t)std::max<int64_t>(n_idx, 1) * (4 + 4 + 27), st, false));  // cls_idx, slot, pt_label, final | events (int4) | nbr
    S1_CUDA(d_polar.reserve((size_t)std::max<int64_t>(n_idx, 1) * 3, st, false));
    S1_CUDA(d_tab.reserve((size_t)std::max<int64_t>(n_tab, 1) * 6, st, false));  // key,min1,min2,coord,kind,label
    S1_CUDA(d_lab.reserve((size_t)std::max<int64_t>(n_lab, 1) * 3, st, false));  // parent,count,first
    S1_CUDA(d_bounds.reserve((size_t)std::max<int64_t>(n_bnd, 1), st, false));
    S1_CUDA(d_pool.reserve((size_t)std::max<int64_t>(n_lab, 1) * 3, st, false));
    S1_CUDA(d_cur.reserve(1, st, false));
    S1_CUDA(d_map.reserve((size_t)std::max<int64_t>(n_lab, 1), st, false));
    B.pts = d_pts; B.labels = d_labels; B.tasks = d_tasks.p; B.ts = d_ts.p;
    B.cls_idx = d_pp.p; B.slot = d_pp.p + n_idx; B.pt_label = d_pp.p + 2 * n_idx; B.final_label = d_pp.p + 3 * n_idx;
    B.events = reinterpret_cast<int4 *>(d_pp.p + 4 * n_idx);  // 4*n_idx ints = 16-byte aligned
    B.nbr = d_pp.p + 8 * n_idx;
    B.polar = d_polar.p;
    B.t_key = d_tab.p; B.t_min1 = d_tab.p + n_tab; B.t_min2 = d_tab.p + 2 * n_tab; B.t_coord = d_tab.p + 3 * n_tab;
    B.t_kind = d_tab.p + 4 * n_tab; B.t_label = d_tab.p + 5 * n_tab;
    B.parent = d_lab.p; B.count = d_lab.p + n_lab; B.first = d_lab.p + 2 * n_lab;
    B.bounds = d_bounds.p; B.pool = d_pool.p; B.pool_cursor = d_cur.p;
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
    k_s1_gather<<<nt, kS1Threads, 0, st>>>(B);
    k_dcvc_prepare<<<nt, kS1Threads, 0, st>>>(B, 0.35, 0.0004, 1.2, 1.2);  // get_json.cpp:205-208
    k_dcvc_replay<<<nt, 32, 0, st>>>(B);
    k_s1_finish<<<nt, kS1Threads, 0, st>>>(B);
    h->launches += 12;
    S1_CUDA(cudaGetLastError());
    S1_CUDA(cudaMemcpyAsync(ts.data(), d_ts.p, nt * sizeof(TaskState), cudaMemcpyDeviceToHost, st));
    S1_CUDA(cudaMemcpyAsync(&cursor, d_cur.p, 8, cudaMemcpyDeviceToHost, st));
    S1_CUDA(cudaStreamSynchronize(st));
    pool.resize((size_t)cursor * 3 + 3);
    if (cursor) S1_CUDA(cudaMemcpyAsync(pool.data(), d_pool.p, (size_t)cursor * 12, cudaMemcpyDeviceToHost, st));
    S1_CUDA(cudaStreamSynchronize(st));
    // ---- instance ids: per scan, tasks in class order; cluster order per policy ----
    int cur_scan = -1, inst_id = 0;
    int64_t node_cursor = 0;
    for (int ti = 0; ti < nt; ++ti) {
      const Task &t = tasks[ti];
      if (t.scan != cur_scan) {
        if (cur_scan >= 0) n_instances[cur_scan] = inst_id;
        for (int s = cur_scan + 1; s <= t.scan; ++s) node_off[s] = node_cursor;
        cur_scan = t.scan; inst_id = 0;
      }
      struct L { int label, count, first; };
      std::vector<L> ls((size_t)ts[ti].ndistinct);
      for (int j = 0; j < ts[ti].ndistinct; ++j) {
        const int *p = &pool[(size_t)(ts[ti].pool_off + j) * 3];
        ls[j] = L{p[0], p[1], p[2]};
      }
      std::vector<L> order;
      if (t.policy == P_DCVC) {
        // labelAnalysis (:394-418): unordered_map keyed by label, filled in ascending point order,
        // emitted in the container's iteration order.
        std::sort(ls.begin(), ls.end(), [](const L &a, const L &b) { return a.first < b.first; });
        std::unordered_map<int, std::vector<int>> label2segIndex;
        std::unordered_map<int, L> info;
        for (const L &l : ls) { label2segIndex[l.label].emplace_back(l.first); info[l.label] = l; }
        for (auto &it : label2segIndex)
          if (info[it.first].count >= t.minSeg) order.push_back(info[it.first]);
      } else if (t.policy == P_GTINST) {
        std::sort(ls.begin(), ls.end(), [](const L &a, const L &b) { return a.label < b.label; });
        for (const L &l : ls) if (l.count > 20) order.push_back(l);  // get_json.cpp:146
      } else {
        order = ls;
      }
      const int mapped = node_map(t.cls);
      for (const L &l : order) {
        InstRec ir{};
        ir.task = ti; ir.label = l.label; ir.inst_id = inst_id++;
        ir.node_slot = -1; ir.node_label = 0;
        if (mapped >= 3 && mapped <= 12) { ir.node_slot = (int)node_cursor++; ir.node_label = (uint32_t)mapped; }
        inst.push_back(ir);
      }
    }
    if (cur_scan >= 0) n_instances[cur_scan] = inst_id;
    for (int s = cur_scan + 1; s <= nscans; ++s) node_off[s] = node_cursor;
    nodes_out.resize((size_t)node_cursor);
    if (!inst.empty()) {
      const int ni = (int)inst.size();
      S1_CUDA(d_inst.reserve(ni, st, false));
      S1_CUDA(cudaMemcpyAsync(d_inst.p, inst.data(), ni * sizeof(InstRec), cudaMemcpyHostToDevice, st));
      S1_CUDA(d_nodes.reserve((size_t)std::max<int64_t>(node_cursor, 1), st, false));
      k_s1_scatter_map<<<(ni + 255) / 256, 256, 0, st>>>(B, d_inst.p, ni, d_map.p);
      if (d_point_instance) k_s1_assign<<<nt, kS1Threads, 0, st>>>(B, d_map.p, d_point_instance);
      k_s1_centroid<<<(ni + 3) / 4, 128, 0, st>>>(B, d_inst.p, ni, d_nodes.p);
      h->launches += 3;
      S1_CUDA(cudaGetLastError());
      if (node_cursor) S1_CUDA(cudaMemcpyAsync(nodes_out.data(), d_nodes.p, (size_t)node_cursor * sizeof(sgtd_node), cudaMemcpyDeviceToHost, st));
      S1_CUDA(cudaStreamSynchronize(st));
    }
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
