// Bucketed WARP-level tensor-core kernels (mma.sync m16n8k8, tf32 in / fp32 accumulate) for the
// TT shapes whose tiles do not fit the tcgen05 kernels' shared-memory budget or tile family
// (ranks 8, 16, 64, 128; any q1; q2 in {4, 8}).  Included by ttb_tt_fast.cu inside its anonymous
// namespace: it reuses the same plan (lookups bucketed by (table, i1), 32-lookup tiles).
//
// Work item = (run of tiles of one bucket) x (j1 block).  For the j1 block the shared B operand is
// B1_j1 = core1[tb][i1][:, j1*r2 : (j1+1)*r2]  (r1 x r2), staged once per run.  With A0 the tile's
// stacked core0 rows (ROWS x r1) and, in the backward, G_j1 = dOut_j1 * C2^T (ROWS x r2):
//   forward   tr0 = A0 * B1_j1            -> out[row][j1][:] = tr0[row][:] * C2_l       (SIMT from fragments)
//   backward  tr0 = A0 * B1_j1 (recompute) -> dCore2 (4-lane transpose + FFMA, red.add)
//             dB  = A0^T * G_j1            -> accumulated in REGISTERS over the run, one red.add pass per run
//             dA  = G_j1 * B1_j1^T         -> red.add into dCore0 rows
// i.e. the reference's per-lookup r1*q1*r2-float atomic scatter (K6/K7) becomes one flush per bucket run.
// Shared-memory strides are padded so the m16n8k8 fragment loads are bank-conflict free in their main role
// (row stride = 4 mod 32 words for A-type tiles, 8 mod 32 for B-type tiles).
#pragma once

namespace bk {

__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                                uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t fbits(float x) { return __float_as_uint(x); }

template <int R1, int R2, int Q2, int ROWS>
struct Cfg {
  static constexpr int kWarps = ROWS / 16;
  static constexpr int kThreads = kWarps * 32;
  static constexpr int kTL = ROWS / 4;   // lookups per tile pass (q0 == 4)
  static constexpr int kSA = R1 + 8;     // words; = 8 mod 32: conflict-free as the A^T operand of dB (the hot
                                         // role); the plain A role is loaded once per tile (hoisted), 2-way
  static constexpr int kSB = R2 + 8;     // words; = 8 mod 32
  static constexpr int kSG = R2 + 8;
  static constexpr int kFwdBytes = (ROWS * kSA + R1 * kSB) * 4 + 1024;
  static constexpr int kBwdBytes = (ROWS * kSA + R1 * kSB + ROWS * kSG + ROWS * Q2) * 4 + 1024;
};

// stage B1_j1 (R1 x R2 block of the core1 slice, row stride n1 floats) -> sB[r][k], tf32
template <int R1, int R2, int THREADS>
__device__ __forceinline__ void stage_b(const float* __restrict__ c1_slice, int n1, int j1, float* sB, int sb,
                                        int tid) {
  constexpr int kVec = R2 / 4;
  for (int it = tid; it < R1 * kVec; it += THREADS) {
    const int r = it / kVec, c4 = it - r * kVec;
    const float4 v = to_tf32(__ldg(reinterpret_cast<const float4*>(c1_slice + (size_t)r * n1 + j1 * R2) + c4));
    *reinterpret_cast<float4*>(sB + r * sb + c4 * 4) = v;
  }
}

// gather A rows: row = l*4 + j0 <- core0[tb][i0_l][j0][0..R1), tf32; rows of padding lookups are zero
template <int R1, int ROWS, int THREADS>
__device__ __forceinline__ void gather_a(const ChainDims& d, const float* __restrict__ core0, int tb,
                                         const LookupRec* srec, int nl, int l_base, float* sA, int sa, int tid) {
  constexpr int kVec = R1 / 4;
  for (int it = tid; it < ROWS * kVec; it += THREADS) {
    const int row = it / kVec, c4 = it - row * kVec;
    const int l = l_base + (row >> 2), j0 = row & 3;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (l < nl)
      v = to_tf32(__ldg(reinterpret_cast<const float4*>(core0 + ((size_t)tb * d.p[0] + srec[l].i0) * d.S[0] + j0 * R1) + c4));
    *reinterpret_cast<float4*>(sA + row * sa + c4 * 4) = v;
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int R1, int R2, int Q2, int ROWS>
__global__ void __launch_bounds__(Cfg<R1, R2, Q2, ROWS>::kThreads)
    tt_fwd_bk_kernel(const ChainDims d, const LookupRec* __restrict__ recs, const int* __restrict__ tile_bucket,
                     const int* __restrict__ tile_begin, const int* __restrict__ tile_count,
                     const int* __restrict__ num_tiles, const CorePtrs cores, float* __restrict__ out) {
  using C = Cfg<R1, R2, Q2, ROWS>;
  extern __shared__ __align__(16) float bk_smem[];
  float* sA = bk_smem;
  float* sB = sA + ROWS * C::kSA;
  LookupRec* srec = reinterpret_cast<LookupRec*>(sB + R1 * C::kSB);  // [32]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q1 = d.q[1], n1 = d.q[1] * R2;
  const int ntiles = *num_tiles;
  for (int item = blockIdx.x; item < ntiles * q1; item += gridDim.x) {
    const int tile = item / q1, j1 = item - tile * q1;
    const int bucket = tile_bucket[tile];
    const int tb = bucket / d.p[1], i1 = bucket - tb * d.p[1];
    const int nl = tile_count[tile];
    __syncthreads();  // previous item's smem readers are done
    if (tid < kTileLookups) {
      LookupRec r;
      r.i0 = 0; r.i2 = 0; r.orow = 0;
      if (tid < nl) r = recs[tile_begin[tile] + tid];
      srec[tid] = r;
    }
    stage_b<R1, R2, C::kThreads>(cores.c[1] + ((size_t)tb * d.p[1] + i1) * d.S[1], n1, j1, sB, C::kSB, tid);
    __syncthreads();
    for (int l_base = 0; l_base < nl; l_base += C::kTL) {  // ROWS == 64: two passes over a 32-lookup tile
      if (l_base) __syncthreads();
      gather_a<R1, ROWS, C::kThreads>(d, cores.c[0], tb, srec, nl, l_base, sA, C::kSA, tid);
      __syncthreads();
      const int m0 = warp * 16;
      const int la = l_base + ((m0 + g) >> 2), lb = l_base + ((m0 + g + 8) >> 2);
      const int j0a = (m0 + g) & 3, j0b = (m0 + g + 8) & 3;
      const bool va = la < nl, vb = lb < nl;
      const float* c2a = cores.c[2] + ((size_t)tb * d.p[2] + srec[va ? la : 0].i2) * d.S[2];
      const float* c2b = cores.c[2] + ((size_t)tb * d.p[2] + srec[vb ? lb : 0].i2) * d.S[2];
      float oa[Q2], ob[Q2];
#pragma unroll
      for (int j = 0; j < Q2; ++j) oa[j] = ob[j] = 0.f;
      // A fragments of this warp's 16 rows do not depend on the n-tile: load them once
      uint32_t af[R1 / 8][4];
#pragma unroll
      for (int ks = 0; ks < R1 / 8; ++ks) {
        const float* pa = sA + (m0 + g) * C::kSA + ks * 8 + t;
        af[ks][0] = fbits(pa[0]);
        af[ks][1] = fbits(pa[8 * C::kSA]);
        af[ks][2] = fbits(pa[4]);
        af[ks][3] = fbits(pa[8 * C::kSA + 4]);
      }
      // core2 rows ka, ka+1 of both lookups for n-tile `ni`, fetched one n-tile ahead of their use
      float4 wa[2][Q2 / 4], wb[2][Q2 / 4];
      auto load_c2 = [&](int ni, float4 (&xa)[2][Q2 / 4], float4 (&xb)[2][Q2 / 4]) {
        const int ka = ni * 8 + 2 * t;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int j4 = 0; j4 < Q2 / 4; ++j4) {
            xa[c][j4] = va ? __ldg(reinterpret_cast<const float4*>(c2a + (ka + c) * Q2) + j4) : make_float4(0.f, 0.f, 0.f, 0.f);
            xb[c][j4] = vb ? __ldg(reinterpret_cast<const float4*>(c2b + (ka + c) * Q2) + j4) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
      };
      load_c2(0, wa, wb);
#pragma unroll 1
      for (int ni = 0; ni < R2 / 8; ++ni) {
        float4 na[2][Q2 / 4], nb2[2][Q2 / 4];
        if (ni + 1 < R2 / 8) load_c2(ni + 1, na, nb2);
        float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < R1 / 8; ++ks) {
          const float* pb = sB + (ks * 8 + t) * C::kSB + ni * 8 + g;
          mma_tf32_16x8x8(c, af[ks][0], af[ks][1], af[ks][2], af[ks][3], fbits(pb[0]), fbits(pb[4 * C::kSB]));
        }
        // last link on the fragment: columns ka = ni*8 + 2t and ka+1 of tr0, rows g (lookup la) and g+8 (lb)
#pragma unroll
        for (int j4 = 0; j4 < Q2 / 4; ++j4) {
          oa[j4 * 4 + 0] = fmaf(c[0], wa[0][j4].x, fmaf(c[1], wa[1][j4].x, oa[j4 * 4 + 0]));
          oa[j4 * 4 + 1] = fmaf(c[0], wa[0][j4].y, fmaf(c[1], wa[1][j4].y, oa[j4 * 4 + 1]));
          oa[j4 * 4 + 2] = fmaf(c[0], wa[0][j4].z, fmaf(c[1], wa[1][j4].z, oa[j4 * 4 + 2]));
          oa[j4 * 4 + 3] = fmaf(c[0], wa[0][j4].w, fmaf(c[1], wa[1][j4].w, oa[j4 * 4 + 3]));
          ob[j4 * 4 + 0] = fmaf(c[2], wb[0][j4].x, fmaf(c[3], wb[1][j4].x, ob[j4 * 4 + 0]));
          ob[j4 * 4 + 1] = fmaf(c[2], wb[0][j4].y, fmaf(c[3], wb[1][j4].y, ob[j4 * 4 + 1]));
          ob[j4 * 4 + 2] = fmaf(c[2], wb[0][j4].z, fmaf(c[3], wb[1][j4].z, ob[j4 * 4 + 2]));
          ob[j4 * 4 + 3] = fmaf(c[2], wb[0][j4].w, fmaf(c[3], wb[1][j4].w, ob[j4 * 4 + 3]));
        }
        if (ni + 1 < R2 / 8) {
#pragma unroll
          for (int c2i = 0; c2i < 2; ++c2i)
#pragma unroll
            for (int j4 = 0; j4 < Q2 / 4; ++j4) {
              wa[c2i][j4] = na[c2i][j4];
              wb[c2i][j4] = nb2[c2i][j4];
            }
        }
      }
      // sum the four column-quarters (lanes t = 0..3 of a row), then lane t == 0 pools into the bag
#pragma unroll
      for (int j = 0; j < Q2; ++j) {
        oa[j] += __shfl_xor_sync(0xffffffffu, oa[j], 1);
        oa[j] += __shfl_xor_sync(0xffffffffu, oa[j], 2);
        ob[j] += __shfl_xor_sync(0xffffffffu, ob[j], 1);
        ob[j] += __shfl_xor_sync(0xffffffffu, ob[j], 2);
      }
      if (t == 0) {
#pragma unroll
        for (int j4 = 0; j4 < Q2; j4 += 4) {
          if (va)
            red_add_f32x4(out + srec[la].orow + (j0a * q1 + j1) * Q2 + j4,
                          make_float4(oa[j4], oa[j4 + 1], oa[j4 + 2], oa[j4 + 3]));
          if (vb)
            red_add_f32x4(out + srec[lb].orow + (j0b * q1 + j1) * Q2 + j4,
                          make_float4(ob[j4], ob[j4 + 1], ob[j4 + 2], ob[j4 + 3]));
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
template <int R1, int R2, int Q2, int ROWS>
__global__ void __launch_bounds__(Cfg<R1, R2, Q2, ROWS>::kThreads)
    tt_bwd_bk_kernel(const ChainDims d, const LookupRec* __restrict__ recs, const int* __restrict__ tile_bucket,
                     const int* __restrict__ tile_begin, const int* __restrict__ tile_count,
                     const int* __restrict__ num_tiles, const int chunk_tiles, const float* __restrict__ d_output,
                     const CorePtrs cores, const CorePtrsRW grads) {
  using C = Cfg<R1, R2, Q2, ROWS>;
  constexpr int kMT = R1 / 16, kNT = R2 / 8;           // dB tiles: kMT x kNT, spread over the warps
  constexpr int kDbTiles = (kMT * kNT + C::kWarps - 1) / C::kWarps;
  extern __shared__ __align__(16) float bk_smem[];
  float* sA = bk_smem;                       // [ROWS][kSA]
  float* sB = sA + ROWS * C::kSA;            // [R1][kSB]
  float* sG = sB + R1 * C::kSB;              // [ROWS][kSG]
  float* sdO = sG + ROWS * C::kSG;           // [ROWS][Q2]  dOut rows of this j1
  LookupRec* srec = reinterpret_cast<LookupRec*>(sdO + ROWS * Q2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q1 = d.q[1], n1 = d.q[1] * R2;
  const int ntiles = *num_tiles;
  const int nchunks = (ntiles + chunk_tiles - 1) / chunk_tiles;

  float db[kDbTiles][4];
  auto flush_db = [&](int bucket, int j1) {
    const int tb = bucket / d.p[1], i1 = bucket - tb * d.p[1];
    float* g1 = grads.c[1] + ((size_t)tb * d.p[1] + i1) * d.S[1] + j1 * R2;
#pragma unroll
    for (int i = 0; i < kDbTiles; ++i) {
      const int tix = warp + i * C::kWarps;
      if (tix < kMT * kNT) {
        const int mi = tix / kNT, ni = tix - mi * kNT;
        float* p0 = g1 + (size_t)(mi * 16 + g) * n1 + ni * 8 + 2 * t;
        red_add_f32x2(p0, db[i][0], db[i][1]);  // columns 2t, 2t+1 are adjacent: one 8-byte reduction
        red_add_f32x2(p0 + (size_t)8 * n1, db[i][2], db[i][3]);
      }
      db[i][0] = db[i][1] = db[i][2] = db[i][3] = 0.f;
    }
  };
#pragma unroll
  for (int i = 0; i < kDbTiles; ++i) db[i][0] = db[i][1] = db[i][2] = db[i][3] = 0.f;

  for (int item = blockIdx.x; item < nchunks * q1; item += gridDim.x) {
    const int chunk = item / q1, j1 = item - chunk * q1;
    const int tile_end = min(ntiles, (chunk + 1) * chunk_tiles);
    int prev_bucket = -1;
    for (int tile = chunk * chunk_tiles; tile < tile_end; ++tile) {
      const int bucket = tile_bucket[tile];
      const int tb = bucket / d.p[1], i1 = bucket - tb * d.p[1];
      const int nl = tile_count[tile];
      const bool new_bucket = bucket != prev_bucket;
      if (new_bucket && prev_bucket >= 0) flush_db(prev_bucket, j1);
      prev_bucket = bucket;
      __syncthreads();
      if (tid < kTileLookups) {
        LookupRec r;
        r.i0 = 0; r.i2 = 0; r.orow = 0;
        if (tid < nl) r = recs[tile_begin[tile] + tid];
        srec[tid] = r;
      }
      if (new_bucket)
        stage_b<R1, R2, C::kThreads>(cores.c[1] + ((size_t)tb * d.p[1] + i1) * d.S[1], n1, j1, sB, C::kSB, tid);
      __syncthreads();
      for (int l_base = 0; l_base < nl; l_base += C::kTL) {
        if (l_base) __syncthreads();
        gather_a<R1, ROWS, C::kThreads>(d, cores.c[0], tb, srec, nl, l_base, sA, C::kSA, tid);
        for (int it = tid; it < ROWS * (Q2 / 4); it += C::kThreads) {  // dOut rows of this j1 block
          const int row = it / (Q2 / 4), c4 = it - row * (Q2 / 4);
          const int l = l_base + (row >> 2), j0 = row & 3;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (l < nl) v = __ldg(reinterpret_cast<const float4*>(d_output + srec[l].orow + (j0 * q1 + j1) * Q2) + c4);
          *reinterpret_cast<float4*>(sdO + row * Q2 + c4 * 4) = v;
        }
        __syncthreads();
        // ---- G[row][k] = sum_j2 dOut[row][j2] * C2_l[k][j2]  (tf32) -> sG
        for (int it = tid; it < ROWS * (R2 / 4); it += C::kThreads) {
          const int row = it / (R2 / 4), k4 = (it - row * (R2 / 4)) * 4;
          const int l = l_base + (row >> 2);
          float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (l < nl) {
            const float* c2 = cores.c[2] + ((size_t)tb * d.p[2] + srec[l].i2) * d.S[2] + k4 * Q2;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j4 = 0; j4 < Q2; j4 += 4) {
              const float4 dv = *reinterpret_cast<const float4*>(sdO + row * Q2 + j4);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(c2 + kk * Q2 + j4));
                acc[kk] = fmaf(dv.x, w.x, fmaf(dv.y, w.y, fmaf(dv.z, w.z, fmaf(dv.w, w.w, acc[kk]))));
              }
            }
            gv = make_float4(to_tf32(acc[0]), to_tf32(acc[1]), to_tf32(acc[2]), to_tf32(acc[3]));
          }
          *reinterpret_cast<float4*>(sG + row * C::kSG + k4) = gv;
        }
        __syncthreads();
        const int m0 = warp * 16;
        const int la = l_base + ((m0 + g) >> 2), lb = l_base + ((m0 + g + 8) >> 2);
        const int j0 = g & 3;  // == (m0 + g) & 3 == (m0 + g + 8) & 3
        // ---- (i) tr0 = A0 * B1_j1 (recompute) and dCore2 from the fragments
        {
        uint32_t af[R1 / 8][4];  // this warp's A fragments are the same for every n-tile: load once
#pragma unroll
        for (int ks = 0; ks < R1 / 8; ++ks) {
          const float* pa = sA + (m0 + g) * C::kSA + ks * 8 + t;
          af[ks][0] = fbits(pa[0]);
          af[ks][1] = fbits(pa[8 * C::kSA]);
          af[ks][2] = fbits(pa[4]);
          af[ks][3] = fbits(pa[8 * C::kSA + 4]);
        }
#pragma unroll 1
        for (int ni = 0; ni < R2 / 8; ++ni) {
          float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int ks = 0; ks < R1 / 8; ++ks) {
            const float* pb = sB + (ks * 8 + t) * C::kSB + ni * 8 + g;
            mma_tf32_16x8x8(c, af[ks][0], af[ks][1], af[ks][2], af[ks][3], fbits(pb[0]), fbits(pb[4 * C::kSB]));
          }
          // 4x4 transpose among the four lanes (j0 = 0..3) of a lookup: lane j0 ends up with output o = j0
          // (o>>1: row half -> lookup la / lb, o&1: column 2t / 2t+1) and the tr0 values of all four rows.
          float w[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int o = j0 ^ k;  // what the partner (j0 ^ k) needs from me is c[its j0] = c[j0 ^ k]
            const float send = (o == 0) ? c[0] : (o == 1) ? c[1] : (o == 2) ? c[2] : c[3];
            w[k] = (k == 0) ? send : __shfl_xor_sync(0xffffffffu, send, k * 4);
          }
          const int o = j0;
          const int l = (o >> 1) ? lb : la;
          if (l < nl) {
            const int row_l0 = ((o >> 1) ? (m0 + g + 8) : (m0 + g)) & ~3;  // first row of lookup l in the tile pass
            float acc[Q2];
#pragma unroll
            for (int j = 0; j < Q2; ++j) acc[j] = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float* dor = sdO + (row_l0 + (j0 ^ k)) * Q2;  // w[k] is tr0 of row j0' = j0 ^ k
#pragma unroll
              for (int j4 = 0; j4 < Q2; j4 += 4) {
                const float4 dv = *reinterpret_cast<const float4*>(dor + j4);
                acc[j4 + 0] = fmaf(w[k], dv.x, acc[j4 + 0]);
                acc[j4 + 1] = fmaf(w[k], dv.y, acc[j4 + 1]);
                acc[j4 + 2] = fmaf(w[k], dv.z, acc[j4 + 2]);
                acc[j4 + 3] = fmaf(w[k], dv.w, acc[j4 + 3]);
              }
            }
            float* g2 = grads.c[2] + ((size_t)tb * d.p[2] + srec[l].i2) * d.S[2] + (ni * 8 + 2 * t + (o & 1)) * Q2;
#pragma unroll
            for (int j4 = 0; j4 < Q2; j4 += 4)
              red_add_f32x4(g2 + j4, make_float4(acc[j4], acc[j4 + 1], acc[j4 + 2], acc[j4 + 3]));
          }
        }
        }
        // ---- (ii) dB[r][k] += sum_rows A0[row][r] * G[row][k]   (registers, over the whole bucket run)
        // k-step outermost: tile tix = warp + i*kWarps has ni = tix % kNT; when kNT divides the warp count
        // every tile of a warp has the same ni, so its G fragment is loaded once per k-step and shared.
#pragma unroll 2
        for (int ks = 0; ks < ROWS / 8; ++ks) {
          constexpr bool kSameNi = (C::kWarps % kNT) == 0;
          uint32_t bs0 = 0, bs1 = 0;
          if (kSameNi) {
            const float* pb = sG + (ks * 8 + t) * C::kSG + (warp % kNT) * 8 + g;
            bs0 = fbits(pb[0]);
            bs1 = fbits(pb[4 * C::kSG]);
          }
#pragma unroll
          for (int i = 0; i < kDbTiles; ++i) {
            const int tix = warp + i * C::kWarps;
            if (tix < kMT * kNT) {
              const int mi = tix / kNT, ni = tix - mi * kNT;
              const float* pa = sA + (ks * 8 + t) * C::kSA + mi * 16 + g;   // A^T(m = r, k = row) = sA[row][r]
              uint32_t b0 = bs0, b1 = bs1;
              if (!kSameNi) {
                const float* pb = sG + (ks * 8 + t) * C::kSG + ni * 8 + g;
                b0 = fbits(pb[0]);
                b1 = fbits(pb[4 * C::kSG]);
              }
              mma_tf32_16x8x8(db[i], fbits(pa[0]), fbits(pa[8]), fbits(pa[4 * C::kSA]), fbits(pa[4 * C::kSA + 8]),
                              b0, b1);
            }
          }
        }
        // ---- (iii) dA[row][r] = sum_k G[row][k] * B1_j1[r][k]  -> dCore0 rows (partial over this j1 block)
        {
          const bool va = la < nl, vb = lb < nl;
          float* g0a = grads.c[0] + ((size_t)tb * d.p[0] + srec[va ? la : 0].i0) * d.S[0] + j0 * R1;
          float* g0b = grads.c[0] + ((size_t)tb * d.p[0] + srec[vb ? lb : 0].i0) * d.S[0] + j0 * R1;
          uint32_t gf[R2 / 8][4];  // G fragments of this warp's rows: independent of the output n-tile
#pragma unroll
          for (int ks = 0; ks < R2 / 8; ++ks) {
            const float* pa = sG + (m0 + g) * C::kSG + ks * 8 + t;
            gf[ks][0] = fbits(pa[0]);
            gf[ks][1] = fbits(pa[8 * C::kSG]);
            gf[ks][2] = fbits(pa[4]);
            gf[ks][3] = fbits(pa[8 * C::kSG + 4]);
          }
#pragma unroll 1
          for (int ni = 0; ni < R1 / 8; ++ni) {
            float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ks = 0; ks < R2 / 8; ++ks) {
              const float* pb = sB + (ni * 8 + g) * C::kSB + ks * 8 + t;    // B^T(k, n = r) = sB[r][k]
              mma_tf32_16x8x8(c, gf[ks][0], gf[ks][1], gf[ks][2], gf[ks][3], fbits(pb[0]), fbits(pb[4]));
            }
            if (va) red_add_f32x2(g0a + ni * 8 + 2 * t, c[0], c[1]);
            if (vb) red_add_f32x2(g0b + ni * 8 + 2 * t, c[2], c[3]);
          }
        }
      }
    }
    if (prev_bucket >= 0) flush_db(prev_bucket, j1);
  }
}

}  // namespace bk
