// database.cu -- stage 3a: GPU-resident descriptor database.
//
// Replaces the std::unordered_map<STDesc_LOC, std::vector<STDesc>> that
// STDescManager::AddSTDescs fills (R/src/STDesc.cpp:149-172,
// R/include/desc/STDesc.h:217-250,370).
//
// HBM layout built by finalize_db():
//   frame store   rec[N] (32 B) + vert[N] (48 B), keyframe-major insertion order;
//                 index g in this order == the reference's (frame, in-frame) order.
//   vote index    entries stable-sorted by the 64-bit key, so a bucket keeps the
//                 reference's insertion order (frame asc, in-frame order):
//                 SoA v_s0/v_s1/v_s2 (f64) + v_frame (u32)  = 28 B per entry,
//                 coalesced for the bucket scans of the vote kernel.
//   key table     open addressing, 16 B headers {key, off, cnt}, load <= 0.5.
//   frame view    per keyframe the (key, g, sides) triples sorted by key (stable), used to
//                 materialise the match list of a selected candidate keyframe.
// The two sorts are cub::DeviceRadixSort (library plumbing, database build is
// not on the query path); every other kernel is hand written.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>

#include "internal.cuh"

namespace sgtd {

__global__ void k_make_keys(const DescRec *rec, int64_t n, uint64_t *key, uint32_t *idx) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  DescRec r = rec[i];
  // position.x = (int)(side_length_[0] + 0.5) ...  (STDesc.cpp:155-161)
  uint32_t x = (uint32_t)__double2int_rz(__dadd_rn(r.s[0], 0.5));
  uint32_t y = (uint32_t)__double2int_rz(__dadd_rn(r.s[1], 0.5));
  uint32_t z = (uint32_t)__double2int_rz(__dadd_rn(r.s[2], 0.5));
  key[i] = pack_key(x, y, z, r.code);
  idx[i] = (uint32_t)i;
}

__global__ void k_gather_index(const DescRec *rec, const uint32_t *perm, int64_t n, uint32_t frame_lo,
                               double *s0, double *s1, double *s2, uint32_t *fr, float4 *pack, uint64_t *pack8) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  DescRec r = rec[perm[i]];
  s0[i] = r.s[0]; s1[i] = r.s[1]; s2[i] = r.s[2];
  fr[i] = r.frame - frame_lo;
  if (pack) pack[i] = make_float4((float)r.s[0], (float)r.s[1], (float)r.s[2], __uint_as_float(r.frame - frame_lo));
  // 8-byte entry of the join: the sides relative to the bucket's cell (cell = (int)(side + 0.5), the key's
  // x, y, z; side - cell lies in [-0.5, 0.5)) in 13-bit fixed point, and the local frame index (25 bits)
  if (!pack8) return;
  uint64_t w = (uint64_t)(r.frame - frame_lo) << kPack8FrameShift;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double cell = (double)__double2int_rz(__dadd_rn(r.s[k], 0.5));
    int u = (int)floor((r.s[k] - cell + 0.5) * (double)(1 << kPack8Bits));
    u = u < 0 ? 0 : (u > (1 << kPack8Bits) - 1 ? (1 << kPack8Bits) - 1 : u);
    w |= (uint64_t)u << (kPack8Bits * k);
  }
  pack8[i] = w;
}

__global__ void k_insert_buckets(const uint64_t *ukeys, const uint32_t *offs, const uint32_t *cnts,
                                 const int *num_runs, Bucket *table, uint64_t mask) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *num_runs) return;
  const uint64_t key = ukeys[i];
  uint64_t pos = mix64(key) & mask;
  while (true) {
    unsigned long long old = atomicCAS((unsigned long long *)&table[pos].key, SGTD_EMPTY_KEY, (unsigned long long)key);
    if (old == SGTD_EMPTY_KEY) { table[pos].off = offs[i]; table[pos].cnt = cnts[i]; return; }
    pos = (pos + 1) & mask;
  }
}

__global__ void k_gather_sides(const DescRec *rec, const uint32_t *perm, int64_t n, double *out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const DescRec r = rec[perm[i]];
  out[3 * i] = r.s[0]; out[3 * i + 1] = r.s[1]; out[3 * i + 2] = r.s[2];
}

__global__ void k_gather_u64(const uint64_t *src, const uint32_t *perm, int64_t n, uint64_t *dst) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[perm[i]];
}

int finalize_db(sgtd_handle *h) {
  if (!h->dirty) return SGTD_OK;
  cudaStream_t st = h->stream;
  const int64_t N = (int64_t)h->rec.n;
  const int64_t F = h->frames_local();
  h->v_cut_parts = 0;
  SGTD_CUDA(h, h->d_frame_off.reserve((size_t)F + 1, st, false));
  h->d_frame_off.n = (size_t)F + 1;
  SGTD_CUDA(h, cudaMemcpyAsync(h->d_frame_off.p, h->frame_off.data(), (size_t)(F + 1) * 8, cudaMemcpyHostToDevice, st));
  if (N == 0) {
    h->n_buckets = 0; h->table_mask = 0;
    SGTD_CUDA(h, h->table.reserve(1, st, false));
    SGTD_CUDA(h, cudaMemsetAsync(h->table.p, 0xFF, sizeof(Bucket), st));
    h->dirty = false;
    return SGTD_OK;
  }
  if (N >= (1ll << 32)) SGTD_FAIL(h, SGTD_E_CAPACITY, "more than 2^32 descriptors on one rank");
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  // scratch: key/idx double buffers, unique keys / counts / offsets, cub temp
  size_t cub1 = 0, cub2 = 0, cub3 = 0, cub4 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub1, (uint64_t *)nullptr, (uint64_t *)nullptr, (uint32_t *)nullptr,
                                  (uint32_t *)nullptr, N, 0, 60, st);
  cub::DeviceRunLengthEncode::Encode(nullptr, cub2, (uint64_t *)nullptr, (uint64_t *)nullptr, (uint32_t *)nullptr,
                                     (int *)nullptr, N, st);
  cub::DeviceScan::ExclusiveSum(nullptr, cub3, (uint32_t *)nullptr, (uint32_t *)nullptr, N, st);
  cub::DeviceRadixSort::SortPairs(nullptr, cub4, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                  (uint32_t *)nullptr, N, 0, 32, st);
  size_t cubb = std::max(std::max(cub1, cub2), std::max(cub3, cub4));
  size_t o = 0;
  size_t o_k0 = o; o += al(N * 8);
  size_t o_k1 = o; o += al(N * 8);
  size_t o_i0 = o; o += al(N * 4);
  size_t o_i1 = o; o += al(N * 4);
  size_t o_uk = o; o += al(N * 8);
  size_t o_cnt = o; o += al(N * 4);
  size_t o_off = o; o += al(N * 4);
  size_t o_nr = o; o += al(16);
  size_t o_cub = o; o += al(cubb);
  SGTD_CUDA(h, h->scratch.reserve(o, st, false));
  unsigned char *S = h->scratch.p;
  uint64_t *k0 = (uint64_t *)(S + o_k0), *k1 = (uint64_t *)(S + o_k1), *uk = (uint64_t *)(S + o_uk);
  uint32_t *i0 = (uint32_t *)(S + o_i0), *i1 = (uint32_t *)(S + o_i1);
  uint32_t *cnt = (uint32_t *)(S + o_cnt), *off = (uint32_t *)(S + o_off);
  int *d_nr = (int *)(S + o_nr);
  const int TB = 256;
  const unsigned GB = (unsigned)((N + TB - 1) / TB);
  k_make_keys<<<GB, TB, 0, st>>>(h->rec.p, N, k0, i0);
  SGTD_LAUNCHED(h);
  SGTD_CUDA(h, cudaGetLastError());
  // stable sort by key: equal keys keep insertion order g  == bucket vector order
  SGTD_CUDA(h, cub::DeviceRadixSort::SortPairs(S + o_cub, cubb, k0, k1, i0, i1, N, 0, 60, st));
  SGTD_LAUNCHED(h);
  SGTD_CUDA(h, h->v_s0.reserve(N, st, false)); SGTD_CUDA(h, h->v_s1.reserve(N, st, false));
  SGTD_CUDA(h, h->v_s2.reserve(N, st, false)); SGTD_CUDA(h, h->v_frame.reserve(N, st, false));
  // the default join (k_vote_join) streams 16-byte float entries; the 8-byte cell-relative entries of the
  // experimental joins (join_impl 0 / 2) are only built on request
  const bool want16 = h->opt.join_impl == 1, want8 = !want16;
  if (want8 && F >= (1ll << (64 - kPack8FrameShift))) SGTD_FAIL(h, SGTD_E_CAPACITY, "more than 2^25 keyframes on one rank");
  if (want16) SGTD_CUDA(h, h->v_pack.reserve(N, st, false));
  if (want8) SGTD_CUDA(h, h->v_pack8.reserve((size_t)N + 4, st, false));
  h->v_s0.n = h->v_s1.n = h->v_s2.n = h->v_frame.n = (size_t)N;
  h->v_pack.n = want16 ? (size_t)N : 0;
  h->v_pack8.n = want8 ? (size_t)N : 0;
  k_gather_index<<<GB, TB, 0, st>>>(h->rec.p, i1, N, (uint32_t)h->frame_lo(), h->v_s0.p, h->v_s1.p, h->v_s2.p,
                                    h->v_frame.p, want16 ? h->v_pack.p : nullptr, want8 ? h->v_pack8.p : nullptr);
  SGTD_LAUNCHED(h);
  SGTD_CUDA(h, cudaGetLastError());
  // buckets
  SGTD_CUDA(h, cub::DeviceRunLengthEncode::Encode(S + o_cub, cubb, k1, uk, cnt, d_nr, N, st));
  SGTD_LAUNCHED(h);
  int nr = 0;
  SGTD_CUDA(h, cudaMemcpyAsync(&nr, d_nr, sizeof(int), cudaMemcpyDeviceToHost, st));
  SGTD_CUDA(h, cudaStreamSynchronize(st));
  SGTD_CUDA(h, cub::DeviceScan::ExclusiveSum(S + o_cub, cubb, cnt, off, nr, st));
  SGTD_LAUNCHED(h);
  uint64_t tsz = 1024;
  while (tsz < 2ull * (uint64_t)nr) tsz <<= 1;
  SGTD_CUDA(h, h->table.reserve(tsz, st, false));
  h->table.n = tsz; h->table_mask = tsz - 1; h->n_buckets = nr;
  SGTD_CUDA(h, cudaMemsetAsync(h->table.p, 0xFF, tsz * sizeof(Bucket), st));
  k_insert_buckets<<<(unsigned)((nr + TB - 1) / TB), TB, 0, st>>>(uk, off, cnt, d_nr, h->table.p, h->table_mask);
  SGTD_LAUNCHED(h);
  SGTD_CUDA(h, cudaGetLastError());
  // frame view: stable sort of the key-sorted sequence by (local) frame
  //   -> order (frame, key, g); per-frame ranges are frame_off[].
  {
    uint32_t *fr_in = cnt;   // reuse: frame of each key-sorted entry
    uint32_t *fr_out = off;
    SGTD_CUDA(h, cudaMemcpyAsync(fr_in, h->v_frame.p, N * 4, cudaMemcpyDeviceToDevice, st));
    int bits = 1;
    while ((1ll << bits) < std::max<int64_t>(F, 2)) ++bits;
    SGTD_CUDA(h, cub::DeviceRadixSort::SortPairs(S + o_cub, cubb, fr_in, fr_out, i1, i0, N, 0, bits, st));
    SGTD_LAUNCHED(h);
    SGTD_CUDA(h, h->f_key.reserve(N, st, false)); SGTD_CUDA(h, h->f_g.reserve(N, st, false));
    h->f_key.n = h->f_g.n = (size_t)N;
    SGTD_CUDA(h, cudaMemcpyAsync(h->f_g.p, i0, N * 4, cudaMemcpyDeviceToDevice, st));
    k_gather_u64<<<GB, TB, 0, st>>>(k0, i0, N, h->f_key.p);  // k0 still holds keys in g order
    SGTD_CUDA(h, h->f_side.reserve((size_t)N * 3, st, false));
    h->f_side.n = (size_t)N * 3;
    k_gather_sides<<<GB, TB, 0, st>>>(h->rec.p, i0, N, h->f_side.p);
    SGTD_LAUNCHED(h);
    SGTD_LAUNCHED(h);
    SGTD_CUDA(h, cudaGetLastError());
  }
  SGTD_CUDA(h, cudaStreamSynchronize(st));
  h->dirty = false;
  return SGTD_OK;
}

}  // namespace sgtd
