// descriptors.cu -- stage 2: triangle descriptor construction on sm_100a.
//
// Replaces STDescManager::BuildSingleScanSTD (R/src/STDesc.cpp:174-315):
//   per node exact kNN (k = descriptor_near_num) over the scan's node centroids,
//   all (m,n) neighbour pairs -> triangle, side-length gate, 3-swap sort,
//   first-come-wins dedup on the mm-truncated side triple, vertex re-ranking,
//   descriptor emission in (i, m, n) order.
//
// Layout: one persistent CTA walks scans; node centroids of the scan live in
// shared memory (float4), kNN is brute force over the smem tile (one anchor per
// thread, 10-entry register insertion list), triangles are enumerated one per
// thread.  "First come wins" is made order-free with a per-CTA open-addressing
// table keyed on the dedup triple that keeps the minimum (i,m,n) sequence
// number (atomicCAS claim + atomicMin); a triangle survives iff it owns the
// minimum.  Pass 1 writes kNN lists + survivor bitmaps + per-scan counts, an
// exclusive scan turns counts into output offsets, pass 2 emits descriptors.
#include <cub/device/device_scan.cuh>

#include "internal.cuh"

namespace sgtd {

constexpr int kBuildThreads = 256;
constexpr int kMaxNear = 16;
constexpr int kMaxNodes = 4096;  // nodes per scan (smem tile + u16 anchor ids)

struct BuildParams {
  const sgtd_node *nodes;
  const int64_t *node_off;  // nscans+1
  const int64_t *word_off;  // nscans+1 : survivor bitmap word offsets
  const uint32_t *frame_ids;
  int nscans, near_num, npairs;
  double min_len, max_len, scale;
  uint16_t *nn;        // [total_nodes * near_num]
  uint32_t *keep;      // survivor bitmap words
  int64_t *counts;     // per scan (+1 slot for the scan)
  // dedup scratch, one region per CTA
  unsigned long long *slot_key;
  uint32_t *slot_seq;
  uint32_t slots_per_cta;  // power of two
};

struct Tri {
  double s[3];   // sorted sides (metres, not yet scaled)
  int v[3];      // which of (p1,p2,p3) is A,B,C
  uint64_t dkey; // packed dedup triple
  bool valid;
};

__device__ __forceinline__ double side_len(const float4 &u, const float4 &v) {
  // float subtraction, then exact double squares, left-to-right sum (STDesc.cpp:198-203)
  double dx = (double)__fsub_rn(u.x, v.x), dy = (double)__fsub_rn(u.y, v.y),
         dz = (double)__fsub_rn(u.z, v.z);
  return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}

__device__ __forceinline__ Tri make_triangle(const float4 &p1, const float4 &p2, const float4 &p3,
                                             double min_len, double max_len) {
  Tri t;
  double a = side_len(p1, p2), b = side_len(p1, p3), c = side_len(p3, p2);
  t.valid = !(a > max_len || b > max_len || c > max_len || a < min_len || b < min_len || c < min_len);
  // each side is tagged with the vertex it does NOT touch: a->p3, b->p2, c->p1.
  // (equivalent to the l1/l2/l3 tag vectors of STDesc.cpp:214-241)
  int la = 2, lb = 1, lc = 0;
  double tmp; int ti;
  if (a > b) { tmp = a; a = b; b = tmp; ti = la; la = lb; lb = ti; }
  if (b > c) { tmp = b; b = c; c = tmp; ti = lb; lb = lc; lc = ti; }
  if (a > b) { tmp = a; a = b; b = tmp; ti = la; la = lb; lb = ti; }
  t.s[0] = a; t.s[1] = b; t.s[2] = c;
  // A = shared by the two shortest sides = vertex opposite the longest, etc. (:253-291)
  t.v[0] = lc; t.v[1] = lb; t.v[2] = la;
  // pcl::PointXYZ d_p (float) = side*1000 ; (int64_t) truncation (:244-248)
  long long kx = __float2ll_rz(__double2float_rn(__dmul_rn(a, 1000.0)));
  long long ky = __float2ll_rz(__double2float_rn(__dmul_rn(b, 1000.0)));
  long long kz = __float2ll_rz(__double2float_rn(__dmul_rn(c, 1000.0)));
  t.dkey = (uint64_t)kx | ((uint64_t)ky << 21) | ((uint64_t)kz << 42);
  return t;
}

__device__ __forceinline__ void pair_from_index(int p, int near_num, int &m, int &n) {
  // p-th pair in the order "for m in 1..near-2, for n in m+1..near-1"
  m = 1;
  int row = near_num - 2;  // pairs with m == 1
  while (p >= row) { p -= row; --row; ++m; }
  n = m + 1 + p;
}

__global__ void __launch_bounds__(kBuildThreads) k_build_pass1(BuildParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4 *s_node = reinterpret_cast<float4 *>(smem_raw);
  __shared__ long long s_count;
  const int tid = threadIdx.x;
  unsigned long long *slot_key = P.slot_key + (size_t)blockIdx.x * P.slots_per_cta;
  uint32_t *slot_seq = P.slot_seq + (size_t)blockIdx.x * P.slots_per_cta;

  for (int s = blockIdx.x; s < P.nscans; s += gridDim.x) {
    const int64_t n0 = P.node_off[s];
    const int K = (int)(P.node_off[s + 1] - n0);
    // fewer nodes than descriptor_near_num (the reference reads stale kNN indices here): the scan keeps
    // its slot in the batch and yields no descriptor (counts[s] stays 0)
    if (K < P.near_num) continue;
    uint16_t *s_nn = reinterpret_cast<uint16_t *>(s_node + K);
    __syncthreads();
    for (int i = tid; i < K; i += kBuildThreads) {
      sgtd_node nd = P.nodes[n0 + i];
      s_node[i] = make_float4(nd.x, nd.y, nd.z, __uint_as_float(nd.label));
    }
    if (tid == 0) s_count = 0;
    const int T = K * P.npairs;
    // table sized for this scan (power of two >= 2T, capped by the region)
    uint32_t tsz = 64;
    while (tsz < 2u * (uint32_t)T && tsz < P.slots_per_cta) tsz <<= 1;
    const uint32_t tmask = tsz - 1;
    for (uint32_t i = tid; i < tsz; i += kBuildThreads) { slot_key[i] = SGTD_EMPTY_KEY; slot_seq[i] = 0xFFFFFFFFu; }
    __syncthreads();
    // ---- exact kNN: FLANN L2_Simple<float> order (dx*dx + dy*dy) + dz*dz, ties -> lower index
    for (int i = tid; i < K; i += kBuildThreads) {
      float dk[kMaxNear]; int ik[kMaxNear];
#pragma unroll
      for (int k = 0; k < kMaxNear; ++k) { dk[k] = __int_as_float(0x7f800000); ik[k] = 0; }
      const float4 q = s_node[i];
      const int last = P.near_num - 1;
      for (int j = 0; j < K; ++j) {
        const float4 p = s_node[j];
        float dx = __fsub_rn(q.x, p.x), dy = __fsub_rn(q.y, p.y), dz = __fsub_rn(q.z, p.z);
        float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        bool ins = false;
#pragma unroll
        for (int k = 0; k < kMaxNear; ++k) if (k == last) ins = d < dk[k];
        if (ins) {
#pragma unroll
          for (int k = 0; k < kMaxNear; ++k) if (k == last) { dk[k] = d; ik[k] = j; }
#pragma unroll
          for (int k = kMaxNear - 1; k >= 1; --k) {
            if (k <= last && dk[k] < dk[k - 1]) {
              float td = dk[k]; dk[k] = dk[k - 1]; dk[k - 1] = td;
              int tj = ik[k]; ik[k] = ik[k - 1]; ik[k - 1] = tj;
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < kMaxNear; ++k)
        if (k < P.near_num) {
          s_nn[i * P.near_num + k] = (uint16_t)ik[k];
          P.nn[(n0 + i) * P.near_num + k] = (uint16_t)ik[k];
        }
    }
    __syncthreads();
    // ---- enumerate triangles, register survivors' minimum sequence number
    for (int t = tid; t < T; t += kBuildThreads) {
      int i = t / P.npairs, m, n;
      pair_from_index(t - i * P.npairs, P.near_num, m, n);
      Tri tr = make_triangle(s_node[i], s_node[s_nn[i * P.near_num + m]], s_node[s_nn[i * P.near_num + n]],
                             P.min_len, P.max_len);
      if (!tr.valid) continue;
      uint32_t pos = (uint32_t)mix64(tr.dkey) & tmask;
      while (true) {
        unsigned long long old = atomicCAS(&slot_key[pos], SGTD_EMPTY_KEY, (unsigned long long)tr.dkey);
        if (old == SGTD_EMPTY_KEY || old == tr.dkey) { atomicMin(&slot_seq[pos], (uint32_t)t); break; }
        pos = (pos + 1) & tmask;
      }
    }
    __syncthreads();
    // ---- survivors: the triangle that owns the minimum sequence number of its key
    const int64_t w0 = P.word_off[s];
    const int rounds = (T + kBuildThreads - 1) / kBuildThreads;
    int my_count = 0;
    for (int r = 0; r < rounds; ++r) {
      int t = r * kBuildThreads + tid;
      bool keep = false;
      if (t < T) {
        int i = t / P.npairs, m, n;
        pair_from_index(t - i * P.npairs, P.near_num, m, n);
        Tri tr = make_triangle(s_node[i], s_node[s_nn[i * P.near_num + m]],
                               s_node[s_nn[i * P.near_num + n]], P.min_len, P.max_len);
        if (tr.valid) {
          uint32_t pos = (uint32_t)mix64(tr.dkey) & tmask;
          // L1 is not coherent with the L2 atomics above: read through L2 (ld.cg)
          while (__ldcg(&slot_key[pos]) != tr.dkey) pos = (pos + 1) & tmask;
          keep = __ldcg(&slot_seq[pos]) == (uint32_t)t;
        }
      }
      uint32_t bal = __ballot_sync(0xffffffffu, keep);
      if ((tid & 31) == 0) {
        if (r * kBuildThreads + tid < T) P.keep[w0 + (r * kBuildThreads + tid) / 32] = bal;
        my_count += __popc(bal);
      }
    }
    if ((tid & 31) == 0 && my_count) atomicAdd((unsigned long long *)&s_count, (unsigned long long)my_count);
    __syncthreads();
    if (tid == 0) P.counts[s] = s_count;
  }
}

struct EmitParams {
  const sgtd_node *nodes;
  const int64_t *node_off, *word_off, *desc_off;
  const uint32_t *frame_ids;
  const uint16_t *nn;
  const uint32_t *keep;
  int nscans, near_num, npairs;
  double min_len, max_len, scale;
  DescRec *rec;
  DescVert *vert;
};

// pass 2: one CTA per scan (persistent), 256 triangles per round, ordered compaction.
__global__ void __launch_bounds__(kBuildThreads) k_build_pass2(EmitParams P) {
  __shared__ int s_warp[kBuildThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int s = blockIdx.x; s < P.nscans; s += gridDim.x) {
    const int64_t n0 = P.node_off[s];
    const int K = (int)(P.node_off[s + 1] - n0);
    const int T = K < P.near_num ? 0 : K * P.npairs;
    const int64_t w0 = P.word_off[s];
    int64_t base = P.desc_off[s];
    const uint32_t frame = P.frame_ids[s];
    const int rounds = (T + kBuildThreads - 1) / kBuildThreads;
    for (int r = 0; r < rounds; ++r) {
      const int t = r * kBuildThreads + tid;
      uint32_t word = 0;
      if (r * kBuildThreads + warp * 32 < T) word = P.keep[w0 + (r * kBuildThreads + warp * 32) / 32];
      const bool keep = (word >> lane) & 1u;
      const int before = __popc(word & ((1u << lane) - 1u));
      __syncthreads();
      if (lane == 0) s_warp[warp] = __popc(word);
      __syncthreads();
      int wbase = 0, total = 0;
#pragma unroll
      for (int w = 0; w < kBuildThreads / 32; ++w) { if (w < warp) wbase += s_warp[w]; total += s_warp[w]; }
      if (keep) {
        int i = t / P.npairs, m, n;
        pair_from_index(t - i * P.npairs, P.near_num, m, n);
        const uint16_t *nn = P.nn + (n0 + i) * P.near_num;
        sgtd_node q[3] = {P.nodes[n0 + i], P.nodes[n0 + nn[m]], P.nodes[n0 + nn[n]]};
        float4 p1 = make_float4(q[0].x, q[0].y, q[0].z, 0.f), p2 = make_float4(q[1].x, q[1].y, q[1].z, 0.f),
               p3 = make_float4(q[2].x, q[2].y, q[2].z, 0.f);
        Tri tr = make_triangle(p1, p2, p3, P.min_len, P.max_len);
        const int64_t o = base + wbase + before;
        DescRec rc;
        rc.s[0] = __dmul_rn(P.scale, tr.s[0]); rc.s[1] = __dmul_rn(P.scale, tr.s[1]); rc.s[2] = __dmul_rn(P.scale, tr.s[2]);
        rc.frame = frame;
        const uint32_t la = q[tr.v[0]].label, lb = q[tr.v[1]].label, lc = q[tr.v[2]].label;
        rc.code = (uint16_t)(((la & 15u) << 8) | ((lb & 15u) << 4) | (lc & 15u));
        rc.pad = 0;
        P.rec[o] = rc;
        DescVert dv;
        const sgtd_node &A = q[tr.v[0]], &B = q[tr.v[1]], &C = q[tr.v[2]];
        dv.a = make_float4(A.x, A.y, A.z, __uint_as_float((uint32_t)i | ((uint32_t)m << 16) | ((uint32_t)n << 24)));
        dv.b = make_float4(B.x, B.y, B.z, __uint_as_float((la & 255u) | ((lb & 255u) << 8) | ((lc & 255u) << 16)));
        dv.c = make_float4(C.x, C.y, C.z, 0.f);
        P.vert[o] = dv;
      }
      base += total;
    }
  }
}

__global__ void k_fill_frame_ids(uint32_t *ids, int n, uint32_t v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ids[i] = v;
}

// Host driver.  d_nodes: device pointer to all nodes; off: host node offsets.
int build_descriptors(sgtd_handle *h, const sgtd_node *d_nodes, const std::vector<int64_t> &off,
                      const std::vector<uint32_t> &frame_ids, sgtd_desc_batch *out) {
  const int nscans = (int)off.size() - 1;
  const int near_num = h->c.near_num;
  if (near_num < 3 || near_num > kMaxNear) SGTD_FAIL(h, SGTD_E_INVALID, "descriptor_near_num must be in [3,16]");
  const int npairs = (near_num - 1) * (near_num - 2) / 2;
  cudaStream_t st = h->stream;
  out->nscans = nscans;
  out->off.assign(nscans + 1, 0);
  out->n = 0;
  if (nscans == 0) return SGTD_OK;
  int Kmax = 0;
  std::vector<int64_t> word_off(nscans + 1, 0);
  for (int s = 0; s < nscans; ++s) {
    int64_t K = off[s + 1] - off[s];
    if (K < near_num) { word_off[s + 1] = word_off[s]; continue; }  // sparse / empty scan: zero descriptors
    if (K > kMaxNodes) SGTD_FAIL(h, SGTD_E_INVALID, "scan has more than 4096 nodes");
    if (K > Kmax) Kmax = (int)K;
    word_off[s + 1] = word_off[s] + (K * npairs + 31) / 32;
  }
  const int64_t total_nodes = off[nscans];
  uint32_t slots = 64;
  while (slots < 2u * (uint32_t)(Kmax * npairs)) slots <<= 1;
  int grid = std::min(nscans, h->sm_count * 4);
  grid = std::max(1, std::min(grid, (int)((64u << 20) / slots)));  // bound the dedup scratch
  // scratch carve-up
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t b_off = 0;
  size_t o_nodeoff = b_off; b_off += al((nscans + 1) * 8);
  size_t o_wordoff = b_off; b_off += al((nscans + 1) * 8);
  size_t o_counts = b_off; b_off += al((nscans + 1) * 8);
  size_t o_descoff = b_off; b_off += al((nscans + 1) * 8);
  size_t o_fid = b_off; b_off += al((size_t)nscans * 4);
  size_t o_nn = b_off; b_off += al((size_t)total_nodes * near_num * 2);
  size_t o_keep = b_off; b_off += al((size_t)word_off[nscans] * 4);
  size_t o_skey = b_off; b_off += al((size_t)grid * slots * 8);
  size_t o_sseq = b_off; b_off += al((size_t)grid * slots * 4);
  size_t cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (int64_t *)nullptr, (int64_t *)nullptr, nscans + 1, st);
  size_t o_cub = b_off; b_off += al(cub_bytes);
  SGTD_CUDA(h, h->scratch.reserve(b_off, st, false));
  unsigned char *S = h->scratch.p;
  SGTD_CUDA(h, cudaMemcpyAsync(S + o_nodeoff, off.data(), (nscans + 1) * 8, cudaMemcpyHostToDevice, st));
  SGTD_CUDA(h, cudaMemcpyAsync(S + o_wordoff, word_off.data(), (nscans + 1) * 8, cudaMemcpyHostToDevice, st));
  SGTD_CUDA(h, cudaMemcpyAsync(S + o_fid, frame_ids.data(), (size_t)nscans * 4, cudaMemcpyHostToDevice, st));
  SGTD_CUDA(h, cudaMemsetAsync(S + o_counts, 0, (nscans + 1) * 8, st));

  BuildParams P{};
  P.nodes = d_nodes; P.node_off = (const int64_t *)(S + o_nodeoff); P.word_off = (const int64_t *)(S + o_wordoff);
  P.frame_ids = (const uint32_t *)(S + o_fid);
  P.nscans = nscans; P.near_num = near_num; P.npairs = npairs;
  P.min_len = h->c.min_len; P.max_len = h->c.max_len; P.scale = h->c.scale;
  P.nn = (uint16_t *)(S + o_nn); P.keep = (uint32_t *)(S + o_keep); P.counts = (int64_t *)(S + o_counts);
  P.slot_key = (unsigned long long *)(S + o_skey); P.slot_seq = (uint32_t *)(S + o_sseq); P.slots_per_cta = slots;
  size_t smem = (size_t)Kmax * (16 + 2 * near_num) + 16;
  if (smem > 48 * 1024)
    SGTD_CUDA(h, cudaFuncSetAttribute(k_build_pass1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_build_pass1<<<grid, kBuildThreads, smem, st>>>(P);
  SGTD_LAUNCHED(h);
  SGTD_CUDA(h, cudaGetLastError());
  cub::DeviceScan::ExclusiveSum(S + o_cub, cub_bytes, (int64_t *)(S + o_counts), (int64_t *)(S + o_descoff), nscans + 1, st);
  SGTD_LAUNCHED(h);
  SGTD_CUDA(h, cudaMemcpyAsync(out->off.data(), S + o_descoff, (nscans + 1) * 8, cudaMemcpyDeviceToHost, st));
  SGTD_CUDA(h, cudaStreamSynchronize(st));
  out->n = out->off[nscans];
  SGTD_CUDA(h, out->rec.reserve((size_t)std::max<int64_t>(out->n, 1), st, false));
  SGTD_CUDA(h, out->vert.reserve((size_t)std::max<int64_t>(out->n, 1), st, false));
  SGTD_CUDA(h, out->d_off.reserve(nscans + 1, st, false));
  out->rec.n = out->vert.n = (size_t)out->n; out->d_off.n = nscans + 1;
  SGTD_CUDA(h, cudaMemcpyAsync(out->d_off.p, S + o_descoff, (nscans + 1) * 8, cudaMemcpyDeviceToDevice, st));
  EmitParams E{};
  E.nodes = d_nodes; E.node_off = P.node_off; E.word_off = P.word_off; E.desc_off = out->d_off.p;
  E.frame_ids = P.frame_ids; E.nn = P.nn; E.keep = P.keep;
  E.nscans = nscans; E.near_num = near_num; E.npairs = npairs;
  E.min_len = P.min_len; E.max_len = P.max_len; E.scale = P.scale;
  E.rec = out->rec.p; E.vert = out->vert.p;
  if (out->n > 0) {
    k_build_pass2<<<grid, kBuildThreads, 0, st>>>(E);
    SGTD_LAUNCHED(h);
    SGTD_CUDA(h, cudaGetLastError());
  }
  return SGTD_OK;
}

}  // namespace sgtd
