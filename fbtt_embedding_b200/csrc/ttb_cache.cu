// LFU software cache in front of the TT table: hash table maintenance, index
// preprocessing (CSR->COO, cache lookup + stable partition) and the gather / scatter
// kernels over uncompressed cached rows.  HBM-bound integer/byte work: coalesced
// 128-bit accesses, one thread per probe for maximum memory-level parallelism,
// warp-aggregated frequency updates.
//
// Semantics follow the reference exactly (integer state is a bit-exact contract):
//   hash            : hashtbl_cuda_utils.cuh:48-76   (MurmurHash3 fmix of lo,hi words; Lemire range)
//   insert / find   : hashtbl_cuda_utils.cuh:102-154 (linear probing, MAX_PROBES = 3,
//                     tt_embeddings_cuda.cu:29; find never stops at an empty slot)
//   populate        : tt_embeddings_cuda.cu:1115-1139, 1260-1336 (stable sort by frequency
//                     descending; cache_state is NOT cleared on eviction)
//   partition       : tt_embeddings_cuda.cu:1436-1479 (TT first in order, cached tail reversed)
#include <cub/device/device_radix_sort.cuh>

#include "ttb_common.cuh"

namespace ttb {

// implemented in ttb_api.cu / ttb_tt_generic.cu
int launch_fwd_generic(const ChainDims&, int64_t, const int64_t*, const int64_t*, const int64_t*,
                       const CorePtrs&, float*, const int32_t* mask, cudaStream_t);

namespace {

constexpr int kMaxProbes = 3;
constexpr long long kUnused = -1;
constexpr int kTile = 1024;  // elements per partition tile (one CTA)

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

// hashtbl_cuda_utils.cuh:48-76
__device__ __forceinline__ uint32_t murmur3_i64(long long key, uint32_t C) {
  const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
  uint32_t h = 0;
  uint32_t k1 = (uint32_t)((unsigned long long)key & 0xffffffffull);
  k1 *= c1;
  k1 = rotl32(k1, 15);
  k1 *= c2;
  h ^= k1;
  h = rotl32(h, 13);
  h = h * 5 + 0xe6546b64u;
  uint32_t k2 = (uint32_t)((unsigned long long)key >> 32);
  k2 *= c1;
  k2 = rotl32(k2, 15);
  k2 *= c2;
  h ^= k2;
  h = rotl32(h, 13);
  h = h * 5 + 0xe6546b64u;
  h ^= 2;
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return (uint32_t)(((unsigned long long)h * (unsigned long long)C) >> 32);
}

// Read the (up to) three probe slots h, h+1, h+2 with two aligned 128-bit loads when the
// window does not wrap; otherwise three scalar loads.
__device__ __forceinline__ void load_probe_window(const long long* __restrict__ tbl, uint32_t h,
                                                  uint32_t C, long long (&k)[kMaxProbes]) {
  if (h + 3 < C) {
    const uint32_t a = h & ~1u;
    const longlong2 v0 = *reinterpret_cast<const longlong2*>(tbl + a);
    const longlong2 v1 = *reinterpret_cast<const longlong2*>(tbl + a + 2);
    if (h & 1u) {
      k[0] = v0.y;
      k[1] = v1.x;
      k[2] = v1.y;
    } else {
      k[0] = v0.x;
      k[1] = v0.y;
      k[2] = v1.x;
    }
  } else {
#pragma unroll
    for (int j = 0; j < kMaxProbes; ++j) k[j] = tbl[(h + j) % C];
  }
}

// hashtbl_find, hashtbl_cuda_utils.cuh:135-154
__device__ __forceinline__ int find_slot(const long long* __restrict__ tbl, uint32_t C,
                                         long long key) {
  const uint32_t h = murmur3_i64(key, C);
  long long k[kMaxProbes];
  load_probe_window(tbl, h, C, k);
#pragma unroll
  for (int j = 0; j < kMaxProbes; ++j) {
    if (k[j] == key) return (int)((h + j) % C);
    if (key == kUnused) return -1;
  }
  return -1;
}

// update_cache_state_kernel + hashtbl_insert<accumulate=true>.  Read-first probing: a slot
// observed holding another key can never become usable inside this kernel (slots only go
// UNUSED -> key here), and observing our own key is what the CAS would return, so the
// atomicCAS is only issued on slots observed empty.  Result identical to the reference's
// CAS-every-slot loop; hot keys cost one vector load pair + one aggregated add.
// the insert of one key: slot it lives in afterwards, or -1 when its probe window is full of other keys
__device__ __forceinline__ int insert_key(long long* __restrict__ tbl, uint32_t C, long long key) {
  const uint32_t h = murmur3_i64(key, C);
  long long k[kMaxProbes];
  load_probe_window(tbl, h, C, k);
#pragma unroll
  for (int j = 0; j < kMaxProbes; ++j) {
    const uint32_t s = (h + j) % C;
    long long seen = k[j];
    if (seen == kUnused) {
      seen = (long long)atomicCAS(reinterpret_cast<unsigned long long*>(tbl + s),
                                  (unsigned long long)kUnused, (unsigned long long)key);
      if (seen == kUnused) seen = key;
    }
    if (seen == key) return (int)s;
  }
  return -1;
}

// warp-aggregate the +1 of lanes that landed on the same slot (zipf traffic); every lane of the warp calls
__device__ __forceinline__ void count_slot(long long* __restrict__ freq, int slot) {
  const unsigned mask = __match_any_sync(0xffffffffu, slot);
  if (slot >= 0) {
    const int leader = __ffs(mask) - 1;
    if ((int)(threadIdx.x & 31) == leader)
      atomicAdd(reinterpret_cast<unsigned long long*>(freq + slot),
                (unsigned long long)__popc(mask));
  }
}

__global__ void __launch_bounds__(256)
    update_cache_state_kernel(const long long nnz, const long long* __restrict__ indices,
                              const uint32_t C, long long* __restrict__ tbl,
                              long long* __restrict__ freq) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = n < nnz;
  const long long key = active ? __ldg(indices + n) : 0;
  const int slot = active ? insert_key(tbl, C, key) : -1;
  count_slot(freq, slot);
}

// Async cache front-end: the steady-state index preprocessing of one step in ONE
// launch and without the reference's device->host round trip (tt_embeddings_cuda.cu:1481-1488).
// Per lookup n: (1) LFU bookkeeping exactly as update_cache_state_kernel (insert + frequency count);
// (2) CSR -> COO row / table of n; (3) loc[n] = cache_state[slot of the key] or -1 -- what
// cache_lookup_kernel (:1356-1375) returns once every insert of the batch has landed: a key lives in
// the slot its own insert returned (inserts never move or delete keys, and a find scans the same
// window in the same order), and cache_state is not written between populates.
// Lookups stay in batch order: there is no partition and no TT count.  The TT kernels take loc as a
// mask (only loc == -1 is theirs), the cache kernels skip loc < 0.
__global__ void __launch_bounds__(256)
    cache_frontend_kernel(const long long nnz, const long long* __restrict__ colidx,
                          const long long num_bags, const int B,
                          const long long* __restrict__ offsets, const uint32_t C,
                          long long* __restrict__ tbl, long long* __restrict__ freq,
                          const int* __restrict__ cache_state, long long* __restrict__ rowidx,
                          long long* __restrict__ tableidx, int* __restrict__ loc) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = n < nnz;
  const long long key = active ? __ldg(colidx + n) : 0;
  const int slot = active ? insert_key(tbl, C, key) : -1;
  count_slot(freq, slot);
  if (!active) return;
  int l = -1;
  if (slot >= 0 && key != kUnused) l = __ldg(cache_state + slot);  // find(-1) is "absent" (hashtbl_cuda_utils.cuh:141)
  loc[n] = l;
  if (n >= __ldg(offsets) && n < __ldg(offsets + num_bags)) {
    const long long b = bag_of_guess(offsets, num_bags, n, nnz);
    rowidx[n] = b % B;
    tableidx[n] = b / B;
  } else {  // not covered by the offsets: no bag to pool into; -2 keeps it out of BOTH paths
    rowidx[n] = 0;  // (the TT kernels take only loc == -1, the cache kernels only loc >= 0)
    tableidx[n] = 0;
    loc[n] = -2;
  }
}

// mark_popular_colidx_kernel, tt_embeddings_cuda.cu:1115-1139
__global__ void __launch_bounds__(256)
    mark_popular_kernel(const long long H, const long long cache_size,
                        long long* __restrict__ sorted_keys, long long* __restrict__ tbl,
                        long long* __restrict__ freq, int* __restrict__ cache_state) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= H) return;
  const long long key = sorted_keys[n];
  if (key != kUnused) {
    const int slot = find_slot(tbl, (uint32_t)H, key);
    if (slot >= 0) {
      if (n < cache_size) {
        cache_state[slot] = (int)n;
      } else {
        tbl[slot] = kUnused;
        freq[slot] = 0;
      }
    }
  } else if (n < cache_size) {
    sorted_keys[n] = 0;  // empty cache line materialises row 0, like the reference (:1135-1138)
  }
}

// CSR -> COO: one thread per nnz entry, binary search of its bag (uniform work for ragged
// bags, fully coalesced stores).  compute_rowidx_kernel, tt_embeddings_cuda.cu:1338-1354.
__global__ void __launch_bounds__(256)
    rowidx_kernel(const long long nnz, const long long num_bags, const int B,
                  const long long* __restrict__ offsets, long long* __restrict__ rowidx,
                  long long* __restrict__ tableidx) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nnz) return;
  if (n < __ldg(offsets) || n >= __ldg(offsets + num_bags)) return;
  // proportional guess + gallop + bisection: one round trip for batches of similar bag lengths instead of
  // log2(bags) dependent loads (8.7 us for 2048 bags in profiles/r2/cfg3_cache_kernels_ncu_summary.txt)
  const long long lo = bag_of_guess(offsets, num_bags, n, nnz);
  rowidx[n] = lo % B;
  tableidx[n] = lo / B;
}

// cache_lookup_kernel (tt_embeddings_cuda.cu:1356-1375) fused with the per-tile TT count.
// loc[n] = cache location, or -1 when the lookup goes to the TT path.
__global__ void __launch_bounds__(kTile)
    cache_lookup_count_kernel(const long long nnz, const long long* __restrict__ colidx,
                              const uint32_t C, const long long* __restrict__ tbl,
                              const int* __restrict__ cache_state, int* __restrict__ loc,
                              int* __restrict__ tile_counts) {
  __shared__ int s_count;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  const long long n = (long long)blockIdx.x * kTile + threadIdx.x;
  bool is_tt = false;
  if (n < nnz) {
    const int slot = find_slot(tbl, C, __ldg(colidx + n));
    int l = -1;
    if (slot != -1) l = __ldg(cache_state + slot);
    is_tt = (l == -1);
    loc[n] = l;
  }
  const unsigned b = __ballot_sync(0xffffffffu, is_tt);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(&s_count, __popc(b));
  __syncthreads();
  if (threadIdx.x == 0) tile_counts[blockIdx.x] = s_count;
}

// stable partition: TT lookups first in order, cached lookups at the tail reversed
// (cub::DevicePartition::Flagged semantics, tt_embeddings_cuda.cu:1436-1479)
__global__ void __launch_bounds__(kTile)
    partition_scatter_kernel(const long long nnz, const long long* __restrict__ colidx,
                             const long long* __restrict__ rowidx, const int* __restrict__ loc,
                             const int* __restrict__ tile_counts, const int num_tiles,
                             long long* __restrict__ out_col, long long* __restrict__ out_row,
                             int* __restrict__ out_loc, int* __restrict__ d_num_tt) {
  __shared__ int s_warp[kTile / 32];
  __shared__ int s_prefix;
  // exclusive prefix of TT counts over preceding tiles (+ grand total on the last tile)
  int part = 0, tot = 0;
  for (int j = threadIdx.x; j < num_tiles; j += kTile) {
    const int c = tile_counts[j];
    tot += c;
    if (j < (int)blockIdx.x) part += c;
  }
  for (int o = 16; o > 0; o >>= 1) {
    part += __shfl_xor_sync(0xffffffffu, part, o);
    tot += __shfl_xor_sync(0xffffffffu, tot, o);
  }
  __shared__ int s_p[kTile / 32], s_t[kTile / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s_p[warp] = part;
    s_t[warp] = tot;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int p = 0, t = 0;
    for (int w = 0; w < kTile / 32; ++w) {
      p += s_p[w];
      t += s_t[w];
    }
    s_prefix = p;
    if (blockIdx.x == 0) *d_num_tt = t;
  }
  const long long n = (long long)blockIdx.x * kTile + threadIdx.x;
  int l = 0;
  bool valid = n < nnz, is_tt = false;
  if (valid) {
    l = loc[n];
    is_tt = (l == -1);
  }
  const unsigned b = __ballot_sync(0xffffffffu, is_tt);
  if (lane == 0) s_warp[warp] = __popc(b);
  __syncthreads();
  int before = 0;  // TT items in earlier warps of this tile
  for (int w = 0; w < warp; ++w) before += s_warp[w];
  const int tt_rank_local = before + __popc(b & ((1u << lane) - 1));
  if (!valid) return;
  const long long tile_start = (long long)blockIdx.x * kTile;
  if (is_tt) {
    const long long dst = (long long)s_prefix + tt_rank_local;
    out_col[dst] = __ldg(colidx + n);
    out_row[dst] = __ldg(rowidx + n);
    out_loc[dst] = l;  // reference leaves the lookup kernel's uninitialised value here; unused
  } else {
    const long long cached_rank = (tile_start - s_prefix) + (threadIdx.x - tt_rank_local);
    const long long dst = nnz - 1 - cached_rank;
    out_col[dst] = __ldg(colidx + n);
    out_row[dst] = __ldg(rowidx + n);
    out_loc[dst] = l;
  }
}

// output[row][:] += cache_weight[loc][:]      (cache_forward_kernel, :1498-1538)
// grad[loc][:]   += alpha * grad_output[row][:]  (cache_backward_sgd/_dense, :1574-1697)
template <bool GATHER>
__global__ void __launch_bounds__(256)
    cache_rows_kernel(const long long nnz, const int D4, const int* __restrict__ loc,
                      const long long* __restrict__ rowidx, const float* __restrict__ src,
                      float* __restrict__ dst, const float alpha) {
  const long long total = nnz * D4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / D4;
    const int c = (int)(i - n * D4);
    const long long l = __ldg(loc + n);
    if (l < 0) continue;  // a TT lookup of an unpartitioned batch (ttb_cache_frontend)
    const long long r = __ldg(rowidx + n);
    if (GATHER) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src) + l * D4 + c);
      red_add_f32x4(dst + (r * D4 + c) * 4, v);
    } else {
      float4 v = __ldg(reinterpret_cast<const float4*>(src) + r * D4 + c);
      v.x *= alpha;
      v.y *= alpha;
      v.z *= alpha;
      v.w *= alpha;
      red_add_f32x4(dst + (l * D4 + c) * 4, v);
    }
  }
}

// cache_backward_rowwise_adagrad_approx_kernel, tt_embeddings_cuda.cu:1735-1795.
// One warp per cached lookup; the weight update is an atomic add of -g*multiplier (the
// reference does a non-atomic read-modify-write that races on duplicate locations).
__global__ void __launch_bounds__(256)
    cache_rowwise_adagrad_kernel(const long long nnz, const int D, const float* __restrict__ go,
                                 const int* __restrict__ loc, const long long* __restrict__ rowidx,
                                 const float lr, const float eps, float* __restrict__ state,
                                 float* __restrict__ w) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int D4 = D >> 2;
  for (long long n = warp; n < nnz; n += nwarps) {
    const long long l = __ldg(loc + n);
    if (l < 0) continue;  // warp-uniform: one warp per lookup
    const long long r = __ldg(rowidx + n);
    const float4* g4 = reinterpret_cast<const float4*>(go + r * D);
    float ss = 0.f;
    for (int c = lane; c < D4; c += 32) {
      const float4 g = __ldg(g4 + c);
      ss += g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w;
    }
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float g2 = ss / D;
    float mult = 0.f;
    if (lane == 0) {
      const float old = atomicAdd(state + l, g2);
      mult = lr * (1.0f / (sqrtf(old + g2) + eps));
    }
    mult = __shfl_sync(0xffffffffu, mult, 0);
    for (int c = lane; c < D4; c += 32) {
      float4 g = __ldg(g4 + c);
      g.x *= -mult;
      g.y *= -mult;
      g.z *= -mult;
      g.w *= -mult;
      red_add_f32x4(w + l * D + c * 4, g);
    }
  }
}

inline unsigned grid_for(long long work, int block, int per_sm = 8) {
  long long g = (work + block - 1) / block;
  const long long cap = (long long)sm_count() * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace
}  // namespace ttb

using namespace ttb;

extern "C" {

int ttb_update_cache_state(int64_t nnz, const int64_t* indices, int64_t hashtbl_size,
                           int64_t* hashtbl, int64_t* cache_freq, cudaStream_t stream) {
  if (nnz == 0) return 0;  // tt_embeddings_cuda.cu:1095-1097
  TTB_CHECK(hashtbl_size > 0, "hashtbl.numel() must be > 0");  // :1099
  TTB_CHECK(hashtbl_size < 2147483647LL, "hashtbl too large");
  TTB_CHECK(indices && hashtbl && cache_freq, "NULL pointer argument");
  const unsigned blocks = (unsigned)((nnz + 255) / 256);
  KernelTimer timer(TTB_KIND_CACHE, stream);
  update_cache_state_kernel<<<blocks, 256, 0, stream>>>(
      nnz, (const long long*)indices, (uint32_t)hashtbl_size, (long long*)hashtbl,
      (long long*)cache_freq);
  TTB_LAUNCH_CHECK();
  return 0;
}

size_t ttb_cache_populate_temp_bytes(int64_t hashtbl_size) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, (const long long*)nullptr,
                                            (long long*)nullptr, (const long long*)nullptr,
                                            (long long*)nullptr, (int)hashtbl_size, 0, 64,
                                            (cudaStream_t)0);
  return bytes + 256;
}

int ttb_cache_populate(const ttb_shape_t* shape, const float* const* cores, int64_t hashtbl_size,
                       int64_t* hashtbl, int64_t* cache_freq, int32_t* cache_state,
                       int64_t cache_size, float* cache_weight, int64_t* sorted_keys,
                       int64_t* sorted_freq, void* temp, size_t temp_bytes, cudaStream_t stream) {
  ChainDims d;
  if (make_chain_dims(shape, &d)) return 1;
  TTB_CHECK(hashtbl_size > 0, "hashtbl.numel() must be > 0");                  // :1271
  TTB_CHECK(hashtbl_size < 2147483647LL, "hashtbl.numel() must be < INT_MAX");  // :1273
  TTB_CHECK(hashtbl_size >= cache_size, "hashtbl.numel() must be >= cache_size");  // :1274
  TTB_CHECK(d.num_tables == 1, "cache requires num_tables == 1 (tt_embeddings_ops.py:458)");
  size_t need = ttb_cache_populate_temp_bytes(hashtbl_size);
  TTB_CHECK(temp && temp_bytes >= need, "cache_populate temp buffer too small (%zu < %zu)",
            temp_bytes, need);
  // K12: stable descending sort of all slots by frequency, carrying the key
  size_t cub_bytes = temp_bytes;
  TTB_CUDA(cub::DeviceRadixSort::SortPairsDescending(
      temp, cub_bytes, (const long long*)cache_freq, (long long*)sorted_freq,
      (const long long*)hashtbl, (long long*)sorted_keys, (int)hashtbl_size, 0, 64, stream));
  count_launch(4);
  // K13
  mark_popular_kernel<<<(unsigned)((hashtbl_size + 255) / 256), 256, 0, stream>>>(
      hashtbl_size, cache_size, (long long*)sorted_keys, (long long*)hashtbl,
      (long long*)cache_freq, cache_state);
  TTB_LAUNCH_CHECK();
  if (cache_size == 0) return 0;
  // K14: materialise the cached rows through the TT chain.  rowidx == NULL means row n,
  // tableidx == NULL means table 0; `+=` into a zero-filled cache_weight is an exact store.
  TTB_CUDA(cudaMemsetAsync(cache_weight, 0, (size_t)cache_size * d.D * sizeof(float), stream));
  ChainDims dd = d;
  dd.B = (int)cache_size;
  CorePtrs c;
  for (int t = 0; t < TTB_MAX_CORES; ++t) c.c[t] = t < d.T ? cores[t] : nullptr;
  return launch_fwd_generic(dd, cache_size, sorted_keys, nullptr, nullptr, c, cache_weight, nullptr, stream);
}

int ttb_preprocess_rowidx(int64_t nnz, int64_t num_bags_total, int32_t B, const int64_t* offsets,
                          int64_t* rowidx, int64_t* tableidx, cudaStream_t stream) {
  if (nnz == 0) return 0;  // :1389-1391
  TTB_CHECK(B > 0 && num_bags_total > 0, "B and number of bags must be > 0");
  TTB_CHECK(offsets && rowidx && tableidx, "NULL pointer argument");
  rowidx_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, stream>>>(
      nnz, num_bags_total, B, (const long long*)offsets, (long long*)rowidx, (long long*)tableidx);
  TTB_LAUNCH_CHECK();
  return 0;
}

int ttb_cache_frontend(int64_t nnz, const int64_t* colidx, int64_t num_bags_total, int32_t B,
                       const int64_t* offsets, int64_t hashtbl_size, int64_t* hashtbl,
                       int64_t* cache_freq, const int32_t* cache_state, int64_t* rowidx,
                       int64_t* tableidx, int32_t* cache_locations, cudaStream_t stream) {
  if (nnz == 0) return 0;
  TTB_CHECK(B > 0 && num_bags_total > 0, "B and number of bags must be > 0");
  TTB_CHECK(hashtbl_size > 0 && hashtbl_size < 2147483647LL, "bad hashtbl size");
  TTB_CHECK(colidx && offsets && hashtbl && cache_freq && cache_state && rowidx && tableidx &&
                cache_locations,
            "NULL pointer argument");
  KernelTimer timer(TTB_KIND_CACHE, stream);
  cache_frontend_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, stream>>>(
      nnz, (const long long*)colidx, num_bags_total, B, (const long long*)offsets,
      (uint32_t)hashtbl_size, (long long*)hashtbl, (long long*)cache_freq, cache_state,
      (long long*)rowidx, (long long*)tableidx, cache_locations);
  TTB_LAUNCH_CHECK();
  return 0;
}

int64_t ttb_preprocess_tile_count(int64_t nnz) { return (nnz + kTile - 1) / kTile + 1 + nnz; }

int ttb_preprocess_cached(int64_t nnz, const int64_t* colidx, const int64_t* rowidx,
                          int64_t hashtbl_size, const int64_t* hashtbl, const int32_t* cache_state,
                          int64_t* out_colidx, int64_t* out_rowidx, int32_t* out_cache_locations,
                          int32_t* tile_scratch, int32_t* h_num_tt, cudaStream_t stream) {
  TTB_CHECK(h_num_tt != nullptr, "h_num_tt is NULL");
  if (nnz == 0) {
    *h_num_tt = 0;
    return 0;
  }
  TTB_CHECK(hashtbl_size > 0 && hashtbl_size < 2147483647LL, "bad hashtbl size");
  TTB_CHECK(colidx && rowidx && hashtbl && cache_state && out_colidx && out_rowidx &&
                out_cache_locations && tile_scratch,
            "NULL pointer argument");
  const int tiles = (int)((nnz + kTile - 1) / kTile);
  KernelTimer timer(TTB_KIND_CACHE, stream);
  int* tile_counts = tile_scratch;
  int* d_num_tt = tile_scratch + tiles;
  int* loc = tile_scratch + tiles + 1;
  cache_lookup_count_kernel<<<tiles, kTile, 0, stream>>>(nnz, (const long long*)colidx,
                                                         (uint32_t)hashtbl_size,
                                                         (const long long*)hashtbl, cache_state,
                                                         loc, tile_counts);
  TTB_LAUNCH_CHECK();
  partition_scatter_kernel<<<tiles, kTile, 0, stream>>>(
      nnz, (const long long*)colidx, (const long long*)rowidx, loc, tile_counts, tiles,
      (long long*)out_colidx, (long long*)out_rowidx, out_cache_locations, d_num_tt);
  TTB_LAUNCH_CHECK();
  // the one device->host control dependency of the op (reference :1481-1488)
  TTB_CUDA(cudaMemcpyAsync(h_num_tt, d_num_tt, sizeof(int), cudaMemcpyDeviceToHost, stream));
  TTB_CUDA(cudaStreamSynchronize(stream));
  return 0;
}

int ttb_cache_forward(int32_t B, int64_t nnz, int32_t D, const int32_t* cache_locations,
                      const int64_t* rowidx, const float* cache_weight, float* output,
                      cudaStream_t stream) {
  TTB_CHECK(B > 0, "B must be > 0");               // :1549
  TTB_CHECK(D > 0 && D % 4 == 0, "D=%d must be > 0 and divisible by 4", D);  // :1551-1552
  if (nnz == 0) return 0;
  TTB_CHECK(cache_locations && rowidx && cache_weight && output, "NULL pointer argument");
  KernelTimer timer(TTB_KIND_CACHE, stream);
  cache_rows_kernel<true><<<grid_for(nnz * (D / 4), 256), 256, 0, stream>>>(
      nnz, D / 4, cache_locations, (const long long*)rowidx, cache_weight, output, 1.0f);
  TTB_LAUNCH_CHECK();
  return 0;
}

int ttb_cache_backward_sgd(int64_t nnz, int32_t D, const float* grad_output,
                           const int32_t* cache_locations, const int64_t* rowidx, float lr,
                           float* cache_weight, cudaStream_t stream) {
  if (nnz == 0) return 0;
  TTB_CHECK(D > 0 && D % 4 == 0, "D=%d must be > 0 and divisible by 4", D);  // :1637-1638
  TTB_CHECK(cache_locations && rowidx && cache_weight && grad_output, "NULL pointer argument");
  KernelTimer timer(TTB_KIND_CACHE, stream);
  cache_rows_kernel<false><<<grid_for(nnz * (D / 4), 256), 256, 0, stream>>>(
      nnz, D / 4, cache_locations, (const long long*)rowidx, grad_output, cache_weight, -lr);
  TTB_LAUNCH_CHECK();
  return 0;
}

int ttb_cache_backward_dense(int64_t nnz, int32_t D, const float* grad_output,
                             const int32_t* cache_locations, const int64_t* rowidx,
                             float* grad_cache_weight, cudaStream_t stream) {
  if (nnz == 0) return 0;
  TTB_CHECK(D > 0 && D % 4 == 0, "D=%d must be > 0 and divisible by 4", D);  // :1714-1715
  TTB_CHECK(cache_locations && rowidx && grad_cache_weight && grad_output, "NULL pointer argument");
  cache_rows_kernel<false><<<grid_for(nnz * (D / 4), 256), 256, 0, stream>>>(
      nnz, D / 4, cache_locations, (const long long*)rowidx, grad_output, grad_cache_weight, 1.0f);
  TTB_LAUNCH_CHECK();
  return 0;
}

int ttb_cache_backward_rowwise_adagrad_approx(int64_t nnz, int32_t D, const float* grad_output,
                                              const int32_t* cache_locations,
                                              const int64_t* rowidx, float lr, float eps,
                                              float* cache_optimizer_state, float* cache_weight,
                                              cudaStream_t stream) {
  if (nnz == 0) return 0;
  TTB_CHECK(D > 0 && D % 4 == 0, "D=%d must be > 0 and divisible by 4", D);  // :1813-1814
  TTB_CHECK(cache_locations && rowidx && cache_weight && grad_output && cache_optimizer_state,
            "NULL pointer argument");
  KernelTimer timer(TTB_KIND_CACHE, stream);
  cache_rowwise_adagrad_kernel<<<grid_for(nnz * 32, 256), 256, 0, stream>>>(
      nnz, D, grad_output, cache_locations, (const long long*)rowidx, lr, eps,
      cache_optimizer_state, cache_weight);
  TTB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
