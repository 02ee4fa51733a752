// tcgen05 TT-EmbeddingBag kernels with 16-bit operands (kind::f16, bf16 inputs, fp32 accumulate in TMEM) for
// equal ranks R in {16, 32, 64, 128}, q0 == 4, q2 in {4, 8}, (q1 * R) % 128 == 0 (% 64 at R = 16).  Included by
// ttb_tt_fast.cu inside its anonymous namespace; consumes the same plan (lookups bucketed by (table, i1), runs of
// <= max_run tiles per bucket).
//
// Why bf16 operands for fp32 cores: every fp32 value is split x = hi + lo (hi = bf16(x), lo = bf16(x - hi)) when it
// is staged, and each product is accumulated as hi*hi + hi*lo + lo*hi in ONE TMEM tile.  Measured on B200
// (tests/cuda/mma_probe3.cu, profiles/r2_first_call/probe3_run.log): max-norm relative error 3.8e-6 against
// 2.5e-4 for the single tf32 MMA the round-1 kernels issued -- fp32-grade results from the tensor pipe.  And 16-bit
// operands have BOTH majors under the standard 128-byte swizzle, so one staged tile in natural row order serves a
// GEMM and its transpose: no transposed copies, every staging store is a 16-byte vector store, the backward's
// tile set shrinks from 211 KB to 96 KB at R = 32 (two CTAs per SM).  With bf16 CORES (BASELINE configs[2]) the
// operands are staged as they are (lo == 0, one term).
//
// Work item = (run of tiles of one bucket) x (128-column block cb of the core-1 slice; 64 columns at R = 16) in the
// backward, (tile) x (block) in the forward.  Per 32-lookup tile (M = 32 lookups x q0 = 128 rows):
//   MMA-1  tr0[128 x 128]  = A0[128 x R] * B1[R x 128 (block cb)]                       forward + recompute
//   SIMT   out[row][j1][:] = tr0[row][j1*R + k] * C2_l[k][:]  (+ bag pooling, red.add)   forward epilogue
//   SIMT   G[128 x 128]    = dOut[row][j1][:] . C2_l[k][:]   (bf16 hi/lo -> smem)        backward
//   MMA-3  dA0[128 x R]    = G * B1^T  (K = 128)            -> red.add into dCore0 rows
//   MMA-2  dB1^T[128 x R] += G^T * A0  (K = 128 rows)        accumulated in TMEM over the run
//   SIMT   dC2_l[k][:]     = sum_rows tr0[row][j1*R + k] * dOut[row][j1][:]  (4-lane reduce-scatter, red.add)
// and at the end of a run the dB1 block is either applied to core 1 straight from TMEM (the run holds the whole
// bucket: SGD / Adagrad on the slice, no gradient scratch, no sweep) or added into the gradient scratch (bucket
// split over several runs; the last of its items to land applies the optimizer to the slice).  The small core-0 /
// core-2 gradients are swept by the last 32 CTAs to finish (they wait only for CTAs that are already running, so
// the launch needs no co-residency) -- the optimizer needs no launch of its own at any BASELINE shape
// (reference: tt_embeddings_cuda.cu:610-649, three dense memsets + three dense sweeps).
#pragma once


namespace xk {

__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn, int b_mn) {
  // kind::f16: c_format [4,6) = 1 (F32); a_format [7,10) = 1 (BF16); b_format [10,13) = 1; a_major bit 15,
  // b_major bit 16 (1 = MN-major); N >> 3 in [17,23); M >> 4 in [24,29)
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// pull a 128-byte line towards the SM without tying up registers (the data is consumed one pipeline phase later)
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// bf16 element (r, col) of a tile X[rows][cols] stored as 64-column (128-byte) blocks with the 128-byte swizzle:
// the K-major image of [MN = rows][K = cols] and the MN-major image of [K = rows][MN = cols] at once
__device__ __forceinline__ uint32_t sw_off(int rows, int r, int col) {
  return (uint32_t)((col >> 6) * rows * 128 + r * 128 + ((((col >> 3) & 7) ^ (r & 7)) << 4) + (col & 7) * 2);
}

// eight fp32 -> eight bf16 hi (+ eight bf16 lo = bf16(x - hi)), 16 bytes each
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hb = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
    const float2 hf = __bfloat1622float2(hb);
    const __nv_bfloat162 lb = __floats2bfloat162_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hb);
    l[i] = *reinterpret_cast<const uint32_t*>(&lb);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// eight consecutive core elements as fp32 (fp32 cores: two 16-byte loads; bf16 cores: one)
__device__ __forceinline__ void load8(const float* __restrict__ p, float (&x)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
  x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* __restrict__ p, float (&x)[8]) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    x[2 * i] = __uint_as_float(w[i] << 16);
    x[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ float load1(const float* p) { return *p; }
__device__ __forceinline__ float load1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void store1(float* p, float v) { *p = v; }
__device__ __forceinline__ void store1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename CoreT>
struct CoreTraits;
template <>
struct CoreTraits<float> {
  static constexpr bool kSplit = true;  // hi + lo operands, three-term products
};
template <>
struct CoreTraits<__nv_bfloat16> {
  static constexpr bool kSplit = false;  // the stored value IS the operand
};

// Phase trace (debug, ttb_trace_set): thread 0 of every CTA stamps %globaltimer into trace[cta*16 + slot]
__device__ __forceinline__ void stamp(long long* trace, int slot) {
  if (trace && threadIdx.x == 0 && slot < 16) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trace[(size_t)blockIdx.x * 16 + slot] = (long long)t;
  }
}

// NB = width of the core-1 column block a work item handles: 128, except at R = 16 where a slice is often only
// q1 * R = 64 columns wide (q1 = 4).  With 64-column blocks dB1^T is an M = 64 accumulator: row n sits in TMEM lane
// (n / 16) * 32 + n % 16 (tests/cuda/mma_probe4.cu).  At R = 64 such blocks shrink the backward's tile set to 80 KB =
// two CTAs per SM -- measured SLOWER (config 5, r = 64: backward 237 us against 163 us; config 4 on one GPU 2.17 ms
// against 1.79 ms): twice the work items, each repeating the A0 gather, MMA-1 and the per-tile synchronisation for half
// the columns.
template <int R>
struct BwdBlock {
  static constexpr int kNB = (R == 16) ? 64 : 128;
};

template <int R, int Q2, int NB = 128>
struct XCfg {
  static_assert(R == 16 || R == 32 || R == 64 || R == 128, "equal ranks 16 / 32 / 64 / 128");
  static_assert(Q2 == 4 || Q2 == 8, "q2 in {4, 8}");
  static_assert((NB == 64 || NB == 128) && NB % R == 0, "column block of 64 or 128 holding whole j1 groups");
  // A0 hi | lo share one 64-column block: hi in columns [0, R), lo in [32, 32 + R); at R = 16 the other columns are
  // zero-filled once per CTA (MMA-2 takes the whole block as its B operand)
  static constexpr bool kPacked = (R <= 32);
  static constexpr int kKS = R / 16;           // K = 16 steps of MMA-1
  static constexpr int kJB = NB / R;           // j1 groups per column block
  static constexpr int kATile = 128 * 128 * ((R + 63) / 64);  // one half (or hi|lo packed) of A0: 16 / 16 / 32 KB
  static constexpr int kABytes = kPacked ? kATile : 2 * kATile;
  static constexpr int kBTile = R * 128 * (NB / 64);   // one half of the B1 block [R rows][NB cols]: 64-column blocks
  static constexpr int kBBytes = 2 * kBTile;
  static constexpr int kGTile = 128 * 128 * (NB / 64);  // one half of G [128 rows][NB cols]
  static constexpr int kGBytes = 2 * kGTile;
  static constexpr int kMeta = 1024;
  // forward epilogue operand: the 32 lookups' core-2 slices (R x Q2 fp32 each) staged in shared memory by the bulk-copy
  // engine when they fit next to A0 / B1 with >= 2 CTAs per SM; read from there a warp's 8 lookups cost ONE wavefront
  // per k (padded stride), from global memory eight (measured: 5 us of a 13 us forward at the README shape)
#ifndef TTB_C2_SMEM_LIMIT_KB
#define TTB_C2_SMEM_LIMIT_KB 100
#endif
  static constexpr int kC2SmemLimit = TTB_C2_SMEM_LIMIT_KB * 1024;
  static constexpr int kC2Stride = R * Q2 + 4;  // floats per lookup; +4: consecutive lookups land 4 banks apart
  static constexpr int kC2Bytes = kTileLookups * kC2Stride * 4;
  static constexpr bool kC2Smem = (kABytes + kBBytes + kC2Bytes) <= kC2SmemLimit;
  static constexpr int kFwdBytes = 1024 + kABytes + kBBytes + (kC2Smem ? kC2Bytes : 0) + kMeta;
  // a forward tile set that fits only once per SM (R = 128: 130 KB) gets 16 warps instead of 8: ncu showed 12.5 % warps
  // active / 13 % issue slots busy with one 256-thread CTA per SM (profiles/r2/cfg5_ranks_8_16_128_ncu_summary.txt)
  static constexpr bool kFwdOnePerSm = kFwdBytes > 113 * 1024;
  static constexpr int kFwdThreads = kFwdOnePerSm ? 512 : 256;
  static constexpr int kBwdBytes = 1024 + kABytes + kBBytes + kGBytes + kMeta;
  static constexpr int kD2Cols = kPacked ? 64 : R;  // packed: D2 = G^T * [A0 hi | A0 lo], the halves are added on read
  static constexpr int kBwdTmem = (NB + R + kD2Cols) <= 256 ? 256 : 512;
  static constexpr bool kBwdTwoPerSm = kBwdBytes <= 113 * 1024 && kBwdTmem <= 256;
  static constexpr int kBwdThreads = kBwdTwoPerSm ? 256 : 512;  // two 256-thread CTAs per SM when the tiles fit twice
  static constexpr int kBwdPerSm = kBwdTwoPerSm ? 2 : 1;
};

struct Meta {
  uint64_t mbar1, mbar2, mbarc;
  uint32_t tmem_slot, pad;
  LookupRec rec[kTileLookups];
};

// ---- staging (generic-proxy 16-byte stores in natural row order) ---------------------------------------------------
// B1 block: core1[slice][r][cb*128 + n] -> XB hi / lo [R rows][128 cols]
template <int R, typename CoreT, int THREADS, int NB = 128>
__device__ __forceinline__ void stage_b1(const CoreT* __restrict__ slice, int n1, int cb, uint8_t* xb, int tid) {
  using C = XCfg<R, 4, NB>;
  constexpr int kChunks = NB / 8;
  for (int u = tid; u < R * kChunks; u += THREADS) {
    const int r = u / kChunks, c8 = u - r * kChunks;
    float x[8];
    load8(slice + (size_t)r * n1 + cb * NB + c8 * 8, x);
    uint4 hi, lo;
    split8(x, hi, lo);
    const uint32_t off = sw_off(R, r, c8 * 8);
    *reinterpret_cast<uint4*>(xb + off) = hi;
    if (CoreTraits<CoreT>::kSplit) *reinterpret_cast<uint4*>(xb + C::kBTile + off) = lo;
  }
}

// A0 rows of the tile: row = l*4 + j0 <- core0[i0_l][j0][0..R); padding lookups are zero rows
template <int R, typename CoreT, int THREADS>
__device__ __forceinline__ void gather_a0(const ChainDims& d, const CoreT* __restrict__ core0, int tb, const Meta* m,
                                          int nl, uint8_t* xa, int tid) {
  using C = XCfg<R, 4>;
  for (int u = tid; u < 128 * (R / 8); u += THREADS) {
    const int row = u / (R / 8), c8 = u - row * (R / 8);
    const int l = row >> 2, j0 = row & 3;
    uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
    if (l < nl) {
      float x[8];
      load8(core0 + ((size_t)tb * d.p[0] + m->rec[l].i0) * d.S[0] + j0 * R + c8 * 8, x);
      split8(x, hi, lo);
    }
    if (C::kPacked) {
      *reinterpret_cast<uint4*>(xa + row * 128 + ((c8 ^ (row & 7)) << 4)) = hi;
      *reinterpret_cast<uint4*>(xa + row * 128 + (((c8 + 4) ^ (row & 7)) << 4)) = lo;  // zero for bf16 cores
    } else {
      const uint32_t off = sw_off(128, row, c8 * 8);
      *reinterpret_cast<uint4*>(xa + off) = hi;
      if (CoreTraits<CoreT>::kSplit) *reinterpret_cast<uint4*>(xa + C::kATile + off) = lo;
    }
  }
}

// R = 16: columns [16, 32) and [48, 64) of the packed A0 block are never written by the gather; MMA-2 reads the whole
// block, so they are zeroed once per CTA (the D2 columns they feed are never read, but they must not hold NaN patterns
// that a debugger / sanitizer run would flag)
template <int BYTES, int THREADS>
__device__ __forceinline__ void zero_a_tile(uint8_t* xa, int tid) {
  for (int u = tid; u < BYTES / 16; u += THREADS) reinterpret_cast<uint4*>(xa)[u] = make_uint4(0u, 0u, 0u, 0u);
}

__device__ __forceinline__ void load_meta(Meta* m, int tid, int nl, const LookupRec* __restrict__ recs) {
  if (tid < kTileLookups) {
    LookupRec r;
    r.i0 = 0;
    r.i2 = 0;
    r.orow = 0;
    if (tid < nl) r = recs[tid];
    m->rec[tid] = r;
  }
}

// ---- MMA issue (one thread) -------------------------------------------------------------------------------------------
// tr0 = A0 * B1(block): A K-major (hi | lo), B MN-major (natural [r][n] rows), three split terms in order of magnitude
template <int R, bool SPLIT, int NB = 128>
__device__ __forceinline__ void issue_mma1(uint32_t d_tmem, const uint8_t* xa, const uint8_t* xb) {
  using C = XCfg<R, 4, NB>;
  constexpr uint32_t kIdesc = make_idesc_bf16(128, NB, 0, 1);
  constexpr int kTerms = SPLIT ? 3 : 1;
  const int ta[3] = {0, 0, 1}, tbh[3] = {0, 1, 0};
  bool first = true;
#pragma unroll
  for (int term = 0; term < kTerms; ++term) {
    const uint32_t abase = smem_u32(xa) + (C::kPacked ? ta[term] * 64 : ta[term] * C::kATile);
    const uint32_t bbase = smem_u32(xb) + tbh[term] * C::kBTile;
#pragma unroll
    for (int ks = 0; ks < C::kKS; ++ks) {
      const uint64_t adesc = make_desc_sw128(abase + (ks >> 2) * (128 * 128) + (ks & 3) * 32, 16, 1024);
      const uint64_t bdesc = make_desc_sw128(bbase + ks * 2048, R * 128, 1024);
      mma_bf16(d_tmem, adesc, bdesc, kIdesc, first ? 0u : 1u);
      first = false;
    }
  }
}

// dA0 = G * B1^T : A = G K-major, B = B1 block K-major ([N = r rows][K = n cols]); K = 128
template <int R, bool SPLIT, int NB = 128>
__device__ __forceinline__ void issue_mma3(uint32_t d_tmem, const uint8_t* xg, const uint8_t* xb) {
  using C = XCfg<R, 4, NB>;
  constexpr uint32_t kIdesc = make_idesc_bf16(128, R, 0, 0);
  constexpr int kTerms = SPLIT ? 3 : 2;  // bf16 cores: G hi * B + G lo * B
  const int tg[3] = {0, SPLIT ? 0 : 1, 1}, tbh[3] = {0, SPLIT ? 1 : 0, 0};
  bool first = true;
#pragma unroll
  for (int term = 0; term < kTerms; ++term) {
    const uint32_t gbase = smem_u32(xg) + tg[term] * C::kGTile;
    const uint32_t bbase = smem_u32(xb) + tbh[term] * C::kBTile;
#pragma unroll
    for (int ks = 0; ks < NB / 16; ++ks) {  // K = NB columns
      const uint64_t adesc = make_desc_sw128(gbase + (ks >> 2) * (128 * 128) + (ks & 3) * 32, 16, 1024);
      const uint64_t bdesc = make_desc_sw128(bbase + (ks >> 2) * (R * 128) + (ks & 3) * 32, 16, 1024);
      mma_bf16(d_tmem, adesc, bdesc, kIdesc, first ? 0u : 1u);
      first = false;
    }
  }
}

// dB1^T (+)= G^T * A0 : A = G MN-major ([K = rows][M = n]), B = A0 MN-major ([K = rows][N = r]); K = 128 rows.
// Packed A0 (R = 32): B is the whole 64-column block [hi | lo], so columns 0-31 hold G^T*A0hi and 32-63 G^T*A0lo.
template <int R, bool SPLIT, int NB = 128>
__device__ __forceinline__ void issue_mma2(uint32_t d_tmem, const uint8_t* xg, const uint8_t* xa, bool accumulate) {
  using C = XCfg<R, 4, NB>;
  constexpr uint32_t kIdesc = make_idesc_bf16(NB, C::kD2Cols, 1, 1);  // M = NB columns of the block (64: see BwdBlock)
  // packed: (G hi, [A hi|A lo]) + (G lo, [A hi|A lo]);  split: (Gh,Ah) (Gh,Al) (Gl,Ah);  bf16 cores: (Gh,A) (Gl,A)
  constexpr int kTerms = C::kPacked ? 2 : (SPLIT ? 3 : 2);
  const int tg[3] = {0, C::kPacked ? 1 : (SPLIT ? 0 : 1), 1}, ta[3] = {0, C::kPacked ? 0 : (SPLIT ? 1 : 0), 0};
  bool first = !accumulate;
#pragma unroll
  for (int term = 0; term < kTerms; ++term) {
    const uint32_t gbase = smem_u32(xg) + tg[term] * C::kGTile;
    const uint32_t abase = smem_u32(xa) + (C::kPacked ? 0 : ta[term] * C::kATile);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint64_t adesc = make_desc_sw128(gbase + ks * 2048, 128 * 128, 1024);
      const uint64_t bdesc = make_desc_sw128(abase + ks * 2048, 128 * 128, 1024);
      mma_bf16(d_tmem, adesc, bdesc, kIdesc, first ? 0u : 1u);
      first = false;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
template <int R, int Q2, typename CoreT>
__global__ void __launch_bounds__(XCfg<R, Q2, BwdBlock<R>::kNB>::kFwdThreads)
    x_fwd_kernel(const ChainDims d, const LookupRec* __restrict__ recs, const int* __restrict__ tile_bucket,
                 const int* __restrict__ tile_begin, const int* __restrict__ tile_count,
                 const int* __restrict__ num_tiles, const CoreT* __restrict__ core0, const CoreT* __restrict__ core1,
                 const CoreT* __restrict__ core2, float* __restrict__ out, long long* trace) {
  constexpr int NB = BwdBlock<R>::kNB;  // columns of the core-1 slice per work item
  using C = XCfg<R, Q2, NB>;
  constexpr int kXFwdThreads = C::kFwdThreads;
  constexpr int HC = NB / (kXFwdThreads / 128);  // columns per thread (2 or 4 warp groups split a block)
  constexpr bool kSplit = CoreTraits<CoreT>::kSplit;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr bool kC2Smem = C::kC2Smem && CoreTraits<CoreT>::kSplit;  // fp32 cores only (raw bulk copies of fp32 slices)
  uint8_t* xa = smem;
  uint8_t* xb = xa + C::kABytes;
  float* sc2 = (float*)(xb + C::kBBytes);
  Meta* meta = (Meta*)(xb + C::kBBytes + (C::kC2Smem ? C::kC2Bytes : 0));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q1 = d.q[1], n1 = q1 * R, ncb = n1 / NB;
  stamp(trace, 0);
  const int nitems = num_tiles[0] * ncb;  // work item = (32-lookup tile, NB-column block)
  // When there are more items than CTAs, a CTA takes the column blocks of one tile in a row, so the tile's metadata,
  // gathered A0 rows and core-2 slices are staged once per tile, not once per block (4 blocks at R = 128).  Tiles
  // stay STRIDED over the CTAs: neighbouring tiles have similar fill (same table, same bucket), and handing a CTA a
  // contiguous range of them was measured 16 % slower on the 26-table batch (forward 0.35 -> 0.41 ms).
  const int group = nitems <= (int)gridDim.x ? 1 : ncb;
  const int ngroups = nitems / group;
  if ((int)blockIdx.x >= ngroups) return;  // whole CTA exits before touching TMEM
  int slot = 2;
  if (warp == 0) tmem_alloc<128>(&meta->tmem_slot);
  if (tid == 0) {
    mbar_init(&meta->mbar1, 1);
    mbar_init(&meta->mbarc, 1);
    fence_mbar_init();
  }
  if (R < 32) zero_a_tile<C::kABytes, kXFwdThreads>(xa, tid);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = meta->tmem_slot;
  uint32_t phase = 0, phasec = 0;
  int prev_tile = -1;
  stamp(trace, 1);

  const int row = (warp & 3) * 32 + lane;  // TMEM lane == tile row (l, j0)
  const int half = warp >> 2;              // columns [HC*half, HC*half + HC) of the block
  const int l = row >> 2, j0 = row & 3;
  const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + half * HC;
  constexpr int kJT = (HC / R) > 0 ? (HC / R) : 1;  // j1 groups inside a thread's columns (R = 16 / 32: 2, else 1)

  for (int n = 0;; ++n) {  // this CTA's n-th item: block n % group of its (n / group)-th tile group
    const int item = ((int)blockIdx.x + (n / group) * (int)gridDim.x) * group + n % group;
    if (item >= nitems) break;
    const int tile = item / ncb, cb = item - tile * ncb;
    const bool fresh = tile != prev_tile;  // false: A0, the metadata and the core-2 slices are still staged
    prev_tile = tile;
    const int bucket = tile_bucket[tile];
    const int tb = bucket / d.p[1];
    const int i1 = bucket - tb * d.p[1];
    const int nl = tile_count[tile];
    {
      if (fresh) load_meta(meta, tid, nl, recs + tile_begin[tile]);
      stage_b1<R, CoreT, kXFwdThreads, NB>(core1 + ((size_t)tb * d.p[1] + i1) * d.S[1], n1, cb, xb, tid);
      __syncthreads();
      stamp(trace, slot++);  // metadata + B1 block staged
      // this thread's output row and its core-2 slice; the part of the slice it will read in the epilogue is
      // prefetched now, so that its latency hides behind the gather and the MMA
      const bool valid = l < nl;
      const CoreT* c2 = core2 + ((size_t)tb * d.p[2] + meta->rec[l].i2) * d.S[2];
      float* orow = out + meta->rec[l].orow + (size_t)j0 * q1 * Q2;
      if (kC2Smem) {
        // the tile's core-2 slices -> shared memory with the bulk-copy engine (lane l copies lookup l's slice); the
        // copies overlap the gather and the MMA, the epilogue waits on the mbarrier
        if (warp == 0 && fresh) {
          constexpr uint32_t kBytes = R * Q2 * 4;
          if (lane == 0) mbar_arrive_expect_tx(&meta->mbarc, (uint32_t)nl * kBytes);
          __syncwarp();
          if (lane < nl)
            tma_bulk_g2s(sc2 + lane * C::kC2Stride,
                         (const float*)core2 + ((size_t)tb * d.p[2] + meta->rec[lane].i2) * d.S[2], kBytes, &meta->mbarc);
        }
      } else if (valid && j0 == 0) {  // one of the lookup's four rows is enough
        const char* pc = reinterpret_cast<const char*>(c2 + (size_t)((half * HC) % R) * Q2);
        constexpr int kBytes = (R < HC ? R : HC) * Q2 * (int)sizeof(CoreT);
#pragma unroll
        for (int b = 0; b < kBytes; b += 128) prefetch_l1(pc + b);
      }
      if (fresh) gather_a0<R, CoreT, kXFwdThreads>(d, core0, tb, meta, nl, xa, tid);
      fence_async_smem();
      tc_fence_before_sync();
      __syncthreads();
      stamp(trace, slot++);  // A0 gathered
      if (tid == 0) {
        tc_fence_after_sync();
        issue_mma1<R, kSplit, NB>(tmem_base, xa, xb);
        mma_commit(&meta->mbar1);
      }
      if (kC2Smem && fresh) {
        mbar_wait(&meta->mbarc, phasec);
        phasec ^= 1;
      }
      mbar_wait(&meta->mbar1, phase);
      phase ^= 1;
      tc_fence_after_sync();
      stamp(trace, slot++);  // MMA done
      float acc[kJT][Q2];
#pragma unroll
      for (int j = 0; j < kJT; ++j)
#pragma unroll
        for (int j2 = 0; j2 < Q2; ++j2) acc[j][j2] = 0.f;
      const bool warp_any = (warp & 3) * 8 < nl;  // a warp of padding lookups has nothing to pool
#pragma unroll
      for (int cc = 0; cc < HC; cc += 16) {
        if (!warp_any) break;
        float v[16];
        tmem_ld16(taddr + cc, v);
        tmem_ld_wait();
        if (valid) {
          const int c = half * HC + cc;        // first column of this chunk inside the block
          const int k0 = c % R;                // rank index of column c (chunks never straddle a j1 group: R >= 16)
          const int jt = (cc / R) < kJT ? (cc / R) : 0;  // which of this thread's j1 groups
#pragma unroll
          for (int k = 0; k < 16; k += 8 / Q2) {
            // one 8-element load covers 8/Q2 consecutive k (Q2 = 4: two k, Q2 = 8: one k)
            float w[8];
            if (kC2Smem) {
              const float4* ps = reinterpret_cast<const float4*>(sc2 + l * C::kC2Stride + (k0 + k) * Q2);
              const float4 a4 = ps[0], b4 = ps[1];
              w[0] = a4.x; w[1] = a4.y; w[2] = a4.z; w[3] = a4.w;
              w[4] = b4.x; w[5] = b4.y; w[6] = b4.z; w[7] = b4.w;
            } else {
              load8(c2 + (size_t)(k0 + k) * Q2, w);
            }
#pragma unroll
            for (int kk = 0; kk < 8 / Q2; ++kk)
#pragma unroll
              for (int j2 = 0; j2 < Q2; ++j2) acc[jt][j2] = fmaf(v[k + kk], w[kk * Q2 + j2], acc[jt][j2]);
          }
        }
      }
      if (valid) {
#pragma unroll
        for (int j = 0; j < kJT; ++j) {
          const int j1 = cb * C::kJB + (half * HC) / R + j;  // global j1 of this thread's j-th group
          float* dst = orow + (size_t)j1 * Q2;
#pragma unroll
          for (int j4 = 0; j4 < Q2; j4 += 4)
            red_add_f32x4(dst + j4, make_float4(acc[j][j4], acc[j][j4 + 1], acc[j][j4 + 2], acc[j][j4 + 3]));
        }
      }
      tc_fence_before_sync();
      __syncthreads();  // A tile, metadata and the TMEM accumulator are reused by the next tile
      stamp(trace, slot++);  // epilogue done
    }
  }
  stamp(trace, 15);
  if (warp == 0) tmem_dealloc<128>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------------
// backward (+ fused optimizer)
// ---------------------------------------------------------------------------------------------------------------------
struct XBwdArgs {
  const LookupRec* recs;
  const int* run_bucket;
  const int* run_begin;
  const int* run_count;
  const int* num_tiles;     // [1] = runs
  const int* bucket_start;  // [nb + 1]
  int* sync_words;          // [3] CTAs that have finished their items
  int* bucket_done;         // [nb] (header, zero on entry / exit): (run, block) items of a SPLIT bucket that have landed
  int nb;
  int tail_sweep;           // 1: the last CTA to finish sweeps the (small) core-0 / core-2 gradients itself
  long long* trace;         // debug phase trace or nullptr
  const float* d_output;
  void* core[3];            // weights (updated in place in the fused modes)
  float* grad[3];           // dense: the op's result; fused: zero-on-entry / zero-on-exit scratch
  float* state[3];          // Adagrad state (fp32) or nullptr
  int optim;                // TTB_OPTIM_*
  float lr, eps;
};

// w -= lr * g  |  s += g*g; w -= lr * g / (sqrt(s) + eps)      (tt_embeddings_cuda.cu:392, 412-414)
template <typename CoreT>
__device__ __forceinline__ void apply_update(CoreT* w, float* s, float g, int optim, float lr, float eps) {
  if (g == 0.f) return;  // untouched element: identical to the dense sweep, which adds 0
  float wv = load1(w);
  if (optim == TTB_OPTIM_ADAGRAD) {
    const float sv = *s + g * g;
    *s = sv;
    wv -= lr * g / (sqrtf(sv) + eps);
  } else {
    wv -= lr * g;
  }
  store1(w, wv);
}

// grid-strided optimizer sweep over one gradient range (float4 granularity; n % 4 == 0 for every TT slice family
// handled here), re-zeroing the scratch
// four consecutive weights as fp32 (16-byte / 8-byte vector access)
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16),
                     __uint_as_float(v.y & 0xffff0000u));
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(__nv_bfloat16* p, float4 v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}
__device__ __forceinline__ float upd(float w, float& s, float g, bool adagrad, float lr, float eps) {
  if (g == 0.f) return w;  // untouched element: identical to the dense sweep, which adds 0
  if (adagrad) {
    s += g * g;
    return w - lr * g / (sqrtf(s) + eps);
  }
  return w - lr * g;
}

// strided optimizer sweep over one gradient range (16-byte granularity; n % 4 == 0 for every TT slice family handled
// here), re-zeroing the scratch.  Gradient, weight and state vectors of U elements are all in flight before the first
// one is looked at: a sweep by few threads is bound by load latency, not bandwidth.
template <typename CoreT>
__device__ __forceinline__ void sweep_range(CoreT* w, float* g, float* s, long long n, int optim, float lr, float eps,
                                            long long first, long long stride) {
  constexpr int U = 4;
  const bool adagrad = optim == TTB_OPTIM_ADAGRAD && s != nullptr;
  const long long n4 = n >> 2;
  for (long long base = first; base < n4; base += stride * U) {
    float4 gv[U], wv[U], sv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = base + u * stride;
      gv[u] = i < n4 ? __ldcg(reinterpret_cast<const float4*>(g) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      wv[u] = i < n4 ? load4(w + (i << 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
      sv[u] = (adagrad && i < n4) ? *(reinterpret_cast<const float4*>(s) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = base + u * stride;
      if (i >= n4 || (gv[u].x == 0.f && gv[u].y == 0.f && gv[u].z == 0.f && gv[u].w == 0.f)) continue;
      wv[u].x = upd(wv[u].x, sv[u].x, gv[u].x, adagrad, lr, eps);
      wv[u].y = upd(wv[u].y, sv[u].y, gv[u].y, adagrad, lr, eps);
      wv[u].z = upd(wv[u].z, sv[u].z, gv[u].z, adagrad, lr, eps);
      wv[u].w = upd(wv[u].w, sv[u].w, gv[u].w, adagrad, lr, eps);
      store4(w + (i << 2), wv[u]);
      if (adagrad) reinterpret_cast<float4*>(s)[i] = sv[u];
      reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

template <int R, int Q2, typename CoreT>
__global__ void __launch_bounds__(XCfg<R, Q2, BwdBlock<R>::kNB>::kBwdThreads, XCfg<R, Q2, BwdBlock<R>::kNB>::kBwdPerSm)
    x_bwd_kernel(const ChainDims d, const XBwdArgs a) {
  constexpr int NB = BwdBlock<R>::kNB;
  using C = XCfg<R, Q2, NB>;
  constexpr bool kSplit = CoreTraits<CoreT>::kSplit;
  constexpr int kThreads = C::kBwdThreads;
  constexpr int KQ = kThreads / 128;  // k-groups: a thread owns k in [kq*KW, (kq+1)*KW) of every j1 group of the block
  constexpr int KW = R / KQ;
  constexpr int NCH = KW / 8;         // 8-wide k chunks per thread
  constexpr int JB = C::kJB;
  constexpr int H = Q2 / 4;
  static_assert(KW % 8 == 0, "a thread's k range is made of 8-wide chunks");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* xa = smem;
  uint8_t* xb = xa + C::kABytes;
  uint8_t* xg = xb + C::kBBytes;
  Meta* meta = (Meta*)(xg + C::kGBytes);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const CoreT* core0 = (const CoreT*)a.core[0];
  const CoreT* core1 = (const CoreT*)a.core[1];
  const CoreT* core2 = (const CoreT*)a.core[2];
  const int q1 = d.q[1], n1 = q1 * R, ncb = n1 / NB;
  const int nitems = a.num_tiles[1] * ncb;
  const bool fused = a.optim != TTB_OPTIM_DENSE;
  const bool has_items = (int)blockIdx.x < nitems;
  uint32_t tmem_base = 0;
  stamp(a.trace, 0);
  int slot = 2;
  if (has_items) {
    if (warp == 0) tmem_alloc<C::kBwdTmem>(&meta->tmem_slot);
    if (tid == 0) {
      mbar_init(&meta->mbar1, 1);
      mbar_init(&meta->mbar2, 2);  // two issuing threads (MMA-3, MMA-2), one commit each
      fence_mbar_init();
    }
    if (R < 32) zero_a_tile<C::kABytes, kThreads>(xa, tid);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    tmem_base = meta->tmem_slot;
  }
  stamp(a.trace, 1);
  const uint32_t tD1 = tmem_base, tD3 = tmem_base + NB, tD2 = tmem_base + NB + R;
  uint32_t phase = 0;

  const int row = (warp & 3) * 32 + lane;  // TMEM lane == tile row (l, j0) == column n of the block in D2
  const int kq = warp >> 2;
  const int l = row >> 2, j0 = row & 3;
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;

  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int run = item / ncb, cb = item - run * ncb;
    const int bucket = a.run_bucket[run];
    const int tb = bucket / d.p[1];
    const int i1 = bucket - tb * d.p[1];
    const int begin = a.run_begin[run], count = a.run_count[run];
    const int bucket_lookups = a.bucket_start[bucket + 1] - a.bucket_start[bucket];
    const size_t slice1 = ((size_t)tb * d.p[1] + i1) * d.S[1];
    stage_b1<R, CoreT, kThreads, NB>(core1 + slice1, n1, cb, xb, tid);
    for (int t0 = 0; t0 < count; t0 += kTileLookups) {
      const int nl = min(kTileLookups, count - t0);
      load_meta(meta, tid, nl, a.recs + begin + t0);
      __syncthreads();
      stamp(a.trace, slot++);  // metadata visible
      const bool valid = l < nl;
      const CoreT* c2 = core2 + ((size_t)tb * d.p[2] + meta->rec[l].i2) * d.S[2];
      if (valid && j0 == 0) {  // the k range this thread's lookup needs in the G phase: latency hides behind the gather
        const char* pc = reinterpret_cast<const char*>(c2 + (size_t)kq * KW * Q2);
#pragma unroll
        for (int b = 0; b < KW * Q2 * (int)sizeof(CoreT); b += 128) prefetch_l1(pc + b);
      }
      gather_a0<R, CoreT, kThreads>(d, core0, tb, meta, nl, xa, tid);
      float4 go[JB][H];  // dOut[l][j0][j1][0..Q2) for the j1 groups of this block
#pragma unroll
      for (int j = 0; j < JB; ++j)
#pragma unroll
        for (int h = 0; h < H; ++h)
          go[j][h] = valid ? __ldg(reinterpret_cast<const float4*>(a.d_output + meta->rec[l].orow +
                                                                  ((size_t)j0 * q1 + cb * JB + j) * Q2 + h * 4))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      fence_async_smem();
      tc_fence_before_sync();
      __syncthreads();
      stamp(a.trace, slot++);  // gather + dOut landed
      if (tid == 0) {
        tc_fence_after_sync();
        issue_mma1<R, kSplit, NB>(tD1, xa, xb);
        mma_commit(&meta->mbar1);
      }
      // ---- G = dOut . C2 while MMA-1 runs: G[row][j*R + k] = sum_j2 dOut[row][j][j2] * C2_l[k][j2]
      // (a warp whose 8 lookups are all padding -- partially filled tile -- only stores its zero rows)
      const bool warp_any = (warp & 3) * 8 < nl;
      if (!warp_any) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
          for (int j = 0; j < JB; ++j) {
            const uint32_t off = sw_off(128, row, j * R + kq * KW + ch * 8);
            *reinterpret_cast<uint4*>(xg + off) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(xg + C::kGTile + off) = make_uint4(0u, 0u, 0u, 0u);
          }
      }
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        if (!warp_any) break;
        const int k0 = kq * KW + ch * 8;
        float4 w[8][H];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (valid) {
            if (Q2 == 4) {
              if ((k & 1) == 0) {
                float x[8];
                load8(c2 + (size_t)(k0 + k) * 4, x);
                w[k][0] = make_float4(x[0], x[1], x[2], x[3]);
                w[k + 1][0] = make_float4(x[4], x[5], x[6], x[7]);
              }
            } else {
              float x[8];
              load8(c2 + (size_t)(k0 + k) * 8, x);
              w[k][0] = make_float4(x[0], x[1], x[2], x[3]);
              w[k][H - 1] = make_float4(x[4], x[5], x[6], x[7]);
            }
          } else {
#pragma unroll
            for (int h = 0; h < H; ++h) w[k][h] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int j = 0; j < JB; ++j) {
          float g[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float acc = 0.f;
#pragma unroll
            for (int h = 0; h < H; ++h)
              acc = fmaf(go[j][h].x, w[k][h].x,
                         fmaf(go[j][h].y, w[k][h].y, fmaf(go[j][h].z, w[k][h].z, fmaf(go[j][h].w, w[k][h].w, acc))));
            g[k] = acc;
          }
          uint4 hi, lo;
          split8(g, hi, lo);
          const uint32_t off = sw_off(128, row, j * R + k0);
          *reinterpret_cast<uint4*>(xg + off) = hi;
          *reinterpret_cast<uint4*>(xg + C::kGTile + off) = lo;
        }
      }
      fence_async_smem();
      tc_fence_before_sync();
      __syncthreads();
      stamp(a.trace, slot++);  // G staged
      // the two gradient GEMMs are issued by two different threads (warps 0 and 1): a single issuer spends ~1 us of
      // descriptor arithmetic + issue per tile in front of its own SIMT share
      if (tid == 0) {
        tc_fence_after_sync();
        issue_mma3<R, kSplit, NB>(tD3, xg, xb);
        mma_commit(&meta->mbar2);
      } else if (tid == 32) {
        tc_fence_after_sync();
        issue_mma2<R, kSplit, NB>(tD2, xg, xa, t0 > 0);
        mma_commit(&meta->mbar2);
      }
      // ---- dC2_l[k][:] += sum_{j0, j} tr0[row][j*R + k] * dOut[row][j][:]   (while MMA-2 / MMA-3 run)
      mbar_wait(&meta->mbar1, phase);
      tc_fence_after_sync();
      // warp-uniform slices (evaluated on the tile's VALID lookups: lane 0's lookup is valid whenever the warp has any)
      const int my_i2 = meta->rec[l].i2, my_i0 = meta->rec[l].i0;
      const int lead_i2 = __shfl_sync(0xffffffffu, my_i2, 0), lead_i0 = __shfl_sync(0xffffffffu, my_i0, 0);
      const bool same_i2 = warp_any && __all_sync(0xffffffffu, !valid || my_i2 == lead_i2);
      const bool same_i0 = warp_any && __all_sync(0xffffffffu, !valid || my_i0 == lead_i0);
      float* g2s = a.grad[2] + ((size_t)tb * d.p[2] + (same_i2 ? lead_i2 : my_i2)) * d.S[2];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        if (!warp_any) break;  // warp-uniform: the TMEM loads below are warp-collective
        const int k0 = kq * KW + ch * 8;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          float4 part[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) part[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          float v[JB][8];  // all of this chunk's TMEM reads in flight before the first wait
#pragma unroll
          for (int j = 0; j < JB; ++j) tmem_ld8(tD1 + lane_addr + j * R + k0, v[j]);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < JB; ++j) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              part[k].x = fmaf(v[j][k], go[j][h].x, part[k].x);
              part[k].y = fmaf(v[j][k], go[j][h].y, part[k].y);
              part[k].z = fmaf(v[j][k], go[j][h].z, part[k].z);
              part[k].w = fmaf(v[j][k], go[j][h].w, part[k].w);
            }
          }
          // reduce-scatter over the 4 rows (j0 = lane & 3) of the lookup: 24 shuffles instead of a 64-shuffle
          // butterfly; lane j0 ends up owning k = k0 + 4*(j0 >> 1) + 2*(j0 & 1) + {0, 1}
          const bool up = (lane & 2) != 0;
          float4 q[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 keep = up ? part[k + 4] : part[k];
            const float4 send = up ? part[k] : part[k + 4];
            q[k].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, 2);
            q[k].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, 2);
            q[k].z = keep.z + __shfl_xor_sync(0xffffffffu, send.z, 2);
            q[k].w = keep.w + __shfl_xor_sync(0xffffffffu, send.w, 2);
          }
          const bool odd = (lane & 1) != 0;
          float4 r2[2];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const float4 keep = odd ? q[k + 2] : q[k];
            const float4 send = odd ? q[k] : q[k + 2];
            r2[k].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, 1);
            r2[k].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, 1);
            r2[k].z = keep.z + __shfl_xor_sync(0xffffffffu, send.z, 1);
            r2[k].w = keep.w + __shfl_xor_sync(0xffffffffu, send.w, 1);
          }
          // Hot slices: when all 8 lookups of this warp hit the SAME core-2 slice (tiny tables, the head of a zipf
          // stream) their contributions are summed in registers first -- 1/8 of the same-address L2 reductions,
          // which serialise in the L2 slice that owns the line.
          bool issue = valid;
          if (same_i2) {
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                r2[k].x += __shfl_xor_sync(0xffffffffu, r2[k].x, o);
                r2[k].y += __shfl_xor_sync(0xffffffffu, r2[k].y, o);
                r2[k].z += __shfl_xor_sync(0xffffffffu, r2[k].z, o);
                r2[k].w += __shfl_xor_sync(0xffffffffu, r2[k].w, o);
              }
            }
            issue = lane < 4;  // padding lookups contributed zeros (their dOut was loaded as zero)
          }
          if (issue) {
            const int kb = k0 + (up ? 4 : 0) + (odd ? 2 : 0);
            red_add_f32x4(g2s + (size_t)(kb + 0) * Q2 + h * 4, r2[0]);
            red_add_f32x4(g2s + (size_t)(kb + 1) * Q2 + h * 4, r2[1]);
          }
        }
      }
      stamp(a.trace, slot++);  // dC2 done
      mbar_wait(&meta->mbar2, phase);
      phase ^= 1;
      tc_fence_after_sync();
      // ---- dCore0[i0_l][j0][r] += dA0[row][r]  (partial over this column block)
      {
        float* g0 = a.grad[0] + ((size_t)tb * d.p[0] + (same_i0 ? lead_i0 : my_i0)) * d.S[0] + j0 * R + kq * KW;
#pragma unroll
        for (int c = 0; c < KW; c += 8) {
          if (!warp_any) break;
          float v[8];
          tmem_ld8(tD3 + lane_addr + kq * KW + c, v);
          tmem_ld_wait();
          bool issue = valid;
          if (same_i0) {  // the warp's 8 lookups share the core-0 slice: sum their rows (same j0) in registers first
            if (!valid) {
#pragma unroll
              for (int k = 0; k < 8; ++k) v[k] = 0.f;  // padding rows of D3 are zero anyway (zero G rows); be explicit
            }
#pragma unroll
            for (int o = 4; o < 32; o <<= 1)
#pragma unroll
              for (int k = 0; k < 8; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
            issue = lane < 4;
          }
          if (issue) {
            red_add_f32x4(g0 + c, make_float4(v[0], v[1], v[2], v[3]));
            red_add_f32x4(g0 + c + 4, make_float4(v[4], v[5], v[6], v[7]));
          }
        }
      }
      tc_fence_before_sync();
      __syncthreads();  // A / G tiles and the metadata are reused by the next tile
    }
    // ---- end of the run: the dB1 block.  D2[n = row][r]; core1 element (r, cb*128 + n).  A run that holds the WHOLE
    // bucket owns the slice: apply the optimizer from TMEM.  Otherwise the partial block goes to the scratch.
    {
      const bool whole = fused && count == bucket_lookups;
      // D2 row n of the block: TMEM lane n for a 128-column block, lane (n / 16) * 32 + n % 16 for a 64-column one
      // (M = 64 accumulator: 16 rows per 32-lane sub-partition, tests/cuda/mma_probe4.cu)
      const int n_blk = (NB == 128) ? row : (warp & 3) * 16 + lane;
      const bool n_ok = (NB == 128) || lane < 16;
      CoreT* w1 = (CoreT*)a.core[1] + slice1 + cb * NB + n_blk;
      float* s1 = a.state[1] ? a.state[1] + slice1 + cb * NB + n_blk : nullptr;
      float* g1 = a.grad[1] + slice1 + cb * NB + n_blk;
#pragma unroll
      for (int c = 0; c < KW; c += 8) {
        float v[8];
        tmem_ld8(tD2 + lane_addr + kq * KW + c, v);
        if (C::kPacked) {
          float v2[8];
          tmem_ld8(tD2 + lane_addr + 32 + kq * KW + c, v2);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] += v2[k];
        } else {
          tmem_ld_wait();
        }
        if (n_ok) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const size_t e = (size_t)(kq * KW + c + k) * n1;
            if (whole)
              apply_update(w1 + e, s1 ? s1 + e : nullptr, v[k], a.optim, a.lr, a.eps);
            else
              red_add_f32(g1 + e, v[k]);
          }
        }
      }
      tc_fence_before_sync();
      if (fused && !whole) {
        // split bucket: the LAST of its (run, block) items to land owns the complete slice in the scratch and applies
        // the optimizer to it (threadfence-reduction pattern: no CTA ever waits for another one)
        fence_gpu();
        __syncthreads();
        if (tid == 0) {
          const int nruns = (bucket_lookups + a.num_tiles[2] * kTileLookups - 1) / (a.num_tiles[2] * kTileLookups);
          const int prev = atomicAdd(a.bucket_done + bucket, 1);
          meta->pad = (prev == nruns * ncb - 1) ? 1u : 0u;
          if (meta->pad) a.bucket_done[bucket] = 0;  // header contract: zero on exit
        }
        __syncthreads();
        if (meta->pad) {
          fence_gpu();
          sweep_range((CoreT*)a.core[1] + slice1, a.grad[1] + slice1, a.state[1] ? a.state[1] + slice1 : nullptr, d.S[1],
                      a.optim, a.lr, a.eps, tid, kThreads);
        }
      }
      __syncthreads();  // the B1 block and D2 are reused by the next item
      stamp(a.trace, slot++);  // run flushed
    }
  }
  stamp(a.trace, 15);
  if (has_items && warp == 0) tmem_dealloc<C::kBwdTmem>(tmem_base);
  if (!fused || !a.tail_sweep) return;

  // ---- cores 0 and 2: swept by the LAST kTail CTAs to finish, each taking an equal share.  Every CTA takes a ticket
  // when its items are done; the holders of the last kTail tickets wait until all tickets are out, i.e. only for
  // CTAs that are RUNNING right now (everything earlier has exited) -- at most kTail SM slots are ever held by
  // waiters, so CTAs that have not been scheduled yet always find a slot: no co-residency requirement, no deadlock.
  constexpr int kTail = 32;
  const int grid = (int)gridDim.x, ktail = grid < kTail ? grid : kTail;
  fence_gpu();
  __syncthreads();
  if (tid == 0) {
    const int ticket = atomicAdd(a.sync_words + 3, 1);
    int share = ticket - (grid - ktail);  // >= 0: one of the last ktail finishers
    if (share >= 0) {
      unsigned spins = 0;
      while (atomicAdd(a.sync_words + 3, 0) < grid) {  // the stragglers are executing: this wait is their tail
        __nanosleep(64);
        if (++spins > (1u << 28)) __trap();
      }
    }
    meta->pad = share >= 0 ? (uint32_t)(share + 1) : 0u;  // 0: not a sweeper
  }
  __syncthreads();
  stamp(a.trace, 12);  // ticket taken (sweepers: every ticket is out)
  if (meta->pad == 0) return;
  fence_gpu();
  // both cores in ONE strided pass, so that their (independent) load -> update -> store chains overlap
  {
    const long long n0 = ((long long)d.num_tables * d.p[0] * d.S[0]) >> 2, n2 = ((long long)d.num_tables * d.p[2] * d.S[2]) >> 2;
    const bool adagrad = a.optim == TTB_OPTIM_ADAGRAD;
    for (long long i = (long long)(meta->pad - 1) * kThreads + tid; i < n0 + n2; i += (long long)ktail * kThreads) {
      const bool first_core = i < n0;
      const long long j = first_core ? i : i - n0;
      CoreT* w = (CoreT*)(first_core ? a.core[0] : a.core[2]) + (j << 2);
      float* g = (first_core ? a.grad[0] : a.grad[2]) + (j << 2);
      float* st = adagrad ? (first_core ? a.state[0] : a.state[2]) + (j << 2) : nullptr;
      const float4 gv = __ldcg(reinterpret_cast<const float4*>(g));
      float4 wv = load4(w);
      float4 sv = adagrad ? *reinterpret_cast<const float4*>(st) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (gv.x == 0.f && gv.y == 0.f && gv.z == 0.f && gv.w == 0.f) continue;
      wv.x = upd(wv.x, sv.x, gv.x, adagrad, a.lr, a.eps);
      wv.y = upd(wv.y, sv.y, gv.y, adagrad, a.lr, a.eps);
      wv.z = upd(wv.z, sv.z, gv.z, adagrad, a.lr, a.eps);
      wv.w = upd(wv.w, sv.w, gv.w, adagrad, a.lr, a.eps);
      store4(w, wv);
      if (adagrad) *reinterpret_cast<float4*>(st) = sv;
      *reinterpret_cast<float4*>(g) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();
  stamp(a.trace, 13);  // swept
  if (tid == 0 && atomicAdd(a.sync_words + 4, 1) == ktail - 1) {  // last sweeper out: the header is zero again
    a.sync_words[3] = 0;
    a.sync_words[4] = 0;
  }
}

// cores 0 and 2 when they are too large for the last CTA of the backward to sweep alone: all CTAs, after the backward
template <typename CoreT>
__global__ void __launch_bounds__(256)
    x_sweep02_kernel(const ChainDims d, void* core0, void* core2, float* g0, float* g2, float* s0, float* s2,
                     const int optim, const float lr, const float eps) {
  const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  sweep_range((CoreT*)core0, g0, s0, (long long)d.num_tables * d.p[0] * d.S[0], optim, lr, eps, first, stride);
  sweep_range((CoreT*)core2, g2, s2, (long long)d.num_tables * d.p[2] * d.S[2], optim, lr, eps, first, stride);
}

}  // namespace xk
