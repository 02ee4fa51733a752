// Generic fp32 (FFMA) TT-EmbeddingBag kernels: any T in 2..4, any p/q/ranks.
//
// These replace the reference's K1+K2+K3 (forward, tt_embeddings_cuda.cu:754-1075)
// and K4..K10 (backward, :79-652) with ONE kernel each plus one optimizer sweep:
// the index decomposition, the chain of small GEMMs, the bag reduction and the
// gradient scatter all happen inside the kernel; no pointer arrays, no `tr`
// round trips through HBM, no chunking by batch_count.
//
// One warp owns one lookup.  Chain state v_t (m_t x r_{t+1}) lives in shared
// memory; TT-core slices are read straight from L2 (the cores of every BASELINE
// config except rank 128 are L2 resident) with 128-bit loads when the slice
// geometry allows.  This is the exact path (fp32 FFMA, rtol 1.3e-6 against the
// reference's tests) and the fallback for shapes the bucketed tensor-core path
// does not cover.
#include "ttb_common.cuh"

namespace ttb {

namespace {

constexpr int kFwdWarps = 4;
constexpr int kBwdWarps = 4;

// out[row][col] = sum_k vin[row][k] * core[k][col]   (row < m, k < K, col < nn)
// STORE: 0 -> vout (shared), 1 -> red.add into global `gout`
template <int MR, bool TO_GLOBAL>
__device__ __forceinline__ void link_scalar(const float* __restrict__ vin,
                                            const float* __restrict__ core, int m, int K, int nn,
                                            float* __restrict__ vout, float* __restrict__ gout,
                                            int lane) {
  const int ntile = (m + MR - 1) / MR;
  const int nitems = ntile * nn;
  for (int item = lane; item < nitems; item += kWarp) {
    const int rt = item / nn;
    const int col = item - rt * nn;
    const int row0 = rt * MR;
    float acc[MR];
#pragma unroll
    for (int i = 0; i < MR; ++i) acc[i] = 0.f;
    const float* vrow = vin + row0 * K;
    if (row0 + MR <= m) {
      for (int k = 0; k < K; ++k) {
        const float c = __ldg(core + (size_t)k * nn + col);
#pragma unroll
        for (int i = 0; i < MR; ++i) acc[i] = fmaf(vrow[i * K + k], c, acc[i]);
      }
    } else {
      for (int k = 0; k < K; ++k) {
        const float c = __ldg(core + (size_t)k * nn + col);
#pragma unroll
        for (int i = 0; i < MR; ++i)
          if (row0 + i < m) acc[i] = fmaf(vrow[i * K + k], c, acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < MR; ++i) {
      if (row0 + i < m) {
        if (TO_GLOBAL)
          red_add_f32(gout + (size_t)(row0 + i) * nn + col, acc[i]);
        else
          vout[(row0 + i) * nn + col] = acc[i];
      }
    }
  }
}

// 4x4 register tile, 128-bit loads of the core slice and of v.  Requires K % 4 == 0,
// nn % 4 == 0, slice base 16-byte aligned, vin/vout 16-byte aligned.
template <bool TO_GLOBAL>
__device__ __forceinline__ void link_vec4(const float* __restrict__ vin,
                                          const float* __restrict__ core, int m, int K, int nn,
                                          float* __restrict__ vout, float* __restrict__ gout,
                                          int lane) {
  const int ncg = nn >> 2;
  const int ntile = (m + 3) >> 2;
  const int nitems = ntile * ncg;
  for (int item = lane; item < nitems; item += kWarp) {
    const int rt = item / ncg;
    const int col = (item - rt * ncg) << 2;
    const int row0 = rt << 2;
    float4 acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < K; k += 4) {
      float4 a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        a[i] = (row0 + i < m) ? *reinterpret_cast<const float4*>(vin + (row0 + i) * K + k)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 c0 = __ldg(reinterpret_cast<const float4*>(core + (size_t)(k + 0) * nn + col));
      const float4 c1 = __ldg(reinterpret_cast<const float4*>(core + (size_t)(k + 1) * nn + col));
      const float4 c2 = __ldg(reinterpret_cast<const float4*>(core + (size_t)(k + 2) * nn + col));
      const float4 c3 = __ldg(reinterpret_cast<const float4*>(core + (size_t)(k + 3) * nn + col));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i].x = fmaf(a[i].x, c0.x, acc[i].x);
        acc[i].y = fmaf(a[i].x, c0.y, acc[i].y);
        acc[i].z = fmaf(a[i].x, c0.z, acc[i].z);
        acc[i].w = fmaf(a[i].x, c0.w, acc[i].w);
        acc[i].x = fmaf(a[i].y, c1.x, acc[i].x);
        acc[i].y = fmaf(a[i].y, c1.y, acc[i].y);
        acc[i].z = fmaf(a[i].y, c1.z, acc[i].z);
        acc[i].w = fmaf(a[i].y, c1.w, acc[i].w);
        acc[i].x = fmaf(a[i].z, c2.x, acc[i].x);
        acc[i].y = fmaf(a[i].z, c2.y, acc[i].y);
        acc[i].z = fmaf(a[i].z, c2.z, acc[i].z);
        acc[i].w = fmaf(a[i].z, c2.w, acc[i].w);
        acc[i].x = fmaf(a[i].w, c3.x, acc[i].x);
        acc[i].y = fmaf(a[i].w, c3.y, acc[i].y);
        acc[i].z = fmaf(a[i].w, c3.z, acc[i].z);
        acc[i].w = fmaf(a[i].w, c3.w, acc[i].w);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (row0 + i < m) {
        if (TO_GLOBAL)
          red_add_f32x4(gout + (size_t)(row0 + i) * nn + col, acc[i]);
        else
          *reinterpret_cast<float4*>(vout + (row0 + i) * nn + col) = acc[i];
      }
    }
  }
}

template <bool TO_GLOBAL>
__device__ __forceinline__ void chain_link(const float* vin, const float* core, int m, int K,
                                           int nn, int S, float* vout, float* gout, int lane) {
  const bool vec_ok = ((K & 3) == 0) && ((nn & 3) == 0) && ((S & 3) == 0);
  if (vec_ok && ((m + 3) >> 2) * (nn >> 2) >= 16) {
    link_vec4<TO_GLOBAL>(vin, core, m, K, nn, vout, gout, lane);
  } else if (((m + 3) >> 2) * nn >= kWarp) {
    link_scalar<4, TO_GLOBAL>(vin, core, m, K, nn, vout, gout, lane);
  } else {
    link_scalar<1, TO_GLOBAL>(vin, core, m, K, nn, vout, gout, lane);
  }
}

struct Digits {
  int i[TTB_MAX_CORES];
  bool ok;
};

__device__ __forceinline__ Digits decompose(const ChainDims& d, long long tb, long long idx) {
  Digits g;
  if (d.het) {
    g.ok = het_digits(d, tb, idx, g.i);
    if (!g.ok) {
#pragma unroll
      for (int t = 0; t < TTB_MAX_CORES; ++t) g.i[t] = 0;
    }
    return g;
  }
  g.ok = idx >= 0;
#pragma unroll
  for (int t = 0; t < TTB_MAX_CORES; ++t) {
    if (t < d.T) {
      const long long qd = idx / d.L[t];
      idx -= qd * d.L[t];
      g.i[t] = (int)qd;
      if (qd >= d.p[t]) g.ok = false;
    } else {
      g.i[t] = 0;
    }
  }
  return g;
}

// -------------------------------------------------------------------------------------
// forward: out[table][row][:] += W[idx]
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFwdWarps* kWarp)
    tt_fwd_generic_kernel(const ChainDims d, const long long nnz,
                          const long long* __restrict__ indices,
                          const long long* __restrict__ rowidx,
                          const long long* __restrict__ tableidx, const CorePtrs cores,
                          float* __restrict__ out, const int* __restrict__ mask) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x / kWarp;
  const int lane = threadIdx.x % kWarp;
  float* buf0 = smem + (size_t)warp * 2 * d.vmax;
  float* buf1 = buf0 + d.vmax;
  for (long long n = (long long)blockIdx.x * kFwdWarps + warp; n < nnz;
       n += (long long)gridDim.x * kFwdWarps) {
    const long long tb = tableidx ? __ldg(tableidx + n) : 0;  // NULL: single table
    const Digits g = decompose(d, tb, __ldg(indices + n));
    const long long row = rowidx ? __ldg(rowidx + n) : n;     // NULL: row n (cache populate)
    if (!g.ok) continue;  // out-of-range index: contributes nothing (reference reads OOB)
    if (mask && __ldg(mask + n) != -1) continue;  // served by the LFU cache (ttb_cache_frontend)
    const long long ctb = d.het ? 0 : tb;  // het: digits are slice numbers in the concatenated cores
    const float* c0 = cores.c[0] + ((size_t)ctb * d.p[0] + g.i[0]) * d.S[0];
    for (int e = lane; e < d.S[0]; e += kWarp) buf0[e] = __ldg(c0 + e);
    __syncwarp();
    float* vin = buf0;
    float* vout = buf1;
    float* orow = out + out_row_offset(d, tb, row);
    for (int t = 1; t < d.T; ++t) {
      const float* ct = cores.c[t] + ((size_t)ctb * d.p[t] + g.i[t]) * d.S[t];
      if (t == d.T - 1)
        chain_link<true>(vin, ct, d.m[t - 1], d.R[t], d.n[t], d.S[t], nullptr, orow, lane);
      else
        chain_link<false>(vin, ct, d.m[t - 1], d.R[t], d.n[t], d.S[t], vout, nullptr, lane);
      __syncwarp();
      float* tmp = vin;
      vin = vout;
      vout = tmp;
    }
  }
}

// -------------------------------------------------------------------------------------
// backward: recompute v_0..v_{T-2}, back-propagate, red.add slice gradients into `grads`
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBwdWarps* kWarp)
    tt_bwd_generic_kernel(const ChainDims d, const long long nnz,
                          const long long* __restrict__ indices,
                          const long long* __restrict__ rowidx,
                          const long long* __restrict__ tableidx,
                          const float* __restrict__ d_output, const CorePtrs cores,
                          const CorePtrsRW grads, const int* __restrict__ mask) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x / kWarp;
  const int lane = threadIdx.x % kWarp;
  const int per_warp = d.vsum + 2 * d.vmax;
  float* vs = smem + (size_t)warp * per_warp;  // v_0 .. v_{T-2}
  float* dvA = vs + d.vsum;
  float* dvB = dvA + d.vmax;
  for (long long n = (long long)blockIdx.x * kBwdWarps + warp; n < nnz;
       n += (long long)gridDim.x * kBwdWarps) {
    const long long tb = __ldg(tableidx + n);
    const Digits g = decompose(d, tb, __ldg(indices + n));
    const long long row = __ldg(rowidx + n);
    if (!g.ok) continue;
    if (mask && __ldg(mask + n) != -1) continue;  // cached lookup: its gradient goes to cache_weight
    const long long ctb = d.het ? 0 : tb;
    // ---- recompute the chain (reference K5, tt_embeddings_cuda.cu:529-545)
    const float* c0 = cores.c[0] + ((size_t)ctb * d.p[0] + g.i[0]) * d.S[0];
    for (int e = lane; e < d.S[0]; e += kWarp) vs[e] = __ldg(c0 + e);
    const float* go = d_output + out_row_offset(d, tb, row);
    for (int e = lane; e < d.D; e += kWarp) dvA[e] = __ldg(go + e);
    __syncwarp();
    for (int t = 1; t < d.T - 1; ++t) {
      const float* ct = cores.c[t] + ((size_t)ctb * d.p[t] + g.i[t]) * d.S[t];
      chain_link<false>(vs + d.voff[t - 1], ct, d.m[t - 1], d.R[t], d.n[t], d.S[t],
                        vs + d.voff[t], nullptr, lane);
      __syncwarp();
    }
    // ---- back-propagate (reference K6/K7/K8, :547-608)
    float* dv = dvA;
    float* dnext = dvB;
    for (int t = d.T - 1; t >= 1; --t) {
      const int m = d.m[t - 1], K = d.R[t], nn = d.n[t];
      const float* prev = vs + d.voff[t - 1];  // [m][K]
      const float* ct = cores.c[t] + ((size_t)ctb * d.p[t] + g.i[t]) * d.S[t];
      float* gt = grads.c[t] + ((size_t)ctb * d.p[t] + g.i[t]) * d.S[t];
      const bool vec = ((nn & 3) == 0) && ((d.S[t] & 3) == 0);
      if (vec) {
        // dCore[k][col..col+3] = sum_row prev[row][k] * dv[row][col..col+3]
        const int ncg = nn >> 2;
        for (int e = lane; e < K * ncg; e += kWarp) {
          const int k = e / ncg;
          const int col = (e - k * ncg) << 2;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int r = 0; r < m; ++r) {
            const float a = prev[r * K + k];
            const float4 b = *reinterpret_cast<const float4*>(dv + r * nn + col);
            acc.x = fmaf(a, b.x, acc.x);
            acc.y = fmaf(a, b.y, acc.y);
            acc.z = fmaf(a, b.z, acc.z);
            acc.w = fmaf(a, b.w, acc.w);
          }
          red_add_f32x4(gt + (size_t)k * nn + col, acc);
        }
        // dPrev[row][k] = sum_col dv[row][col] * core[k][col]
        for (int e = lane; e < m * K; e += kWarp) {
          const int r = e / K;
          const int k = e - r * K;
          float acc = 0.f;
          const float* crow = ct + (size_t)k * nn;
          const float* drow = dv + r * nn;
          for (int col = 0; col < nn; col += 4) {
            const float4 c = __ldg(reinterpret_cast<const float4*>(crow + col));
            const float4 b = *reinterpret_cast<const float4*>(drow + col);
            acc = fmaf(b.x, c.x, acc);
            acc = fmaf(b.y, c.y, acc);
            acc = fmaf(b.z, c.z, acc);
            acc = fmaf(b.w, c.w, acc);
          }
          dnext[e] = acc;
        }
      } else {
        for (int e = lane; e < K * nn; e += kWarp) {
          const int k = e / nn;
          const int col = e - k * nn;
          float acc = 0.f;
          for (int r = 0; r < m; ++r) acc = fmaf(prev[r * K + k], dv[r * nn + col], acc);
          red_add_f32(gt + e, acc);
        }
        for (int e = lane; e < m * K; e += kWarp) {
          const int r = e / K;
          const int k = e - r * K;
          float acc = 0.f;
          const float* crow = ct + (size_t)k * nn;
          const float* drow = dv + r * nn;
          for (int col = 0; col < nn; ++col) acc = fmaf(drow[col], __ldg(crow + col), acc);
          dnext[e] = acc;
        }
      }
      __syncwarp();
      float* tmp = dv;
      dv = dnext;
      dnext = tmp;
    }
    float* g0 = grads.c[0] + ((size_t)ctb * d.p[0] + g.i[0]) * d.S[0];
    for (int e = lane; e < d.S[0]; e += kWarp) red_add_f32(g0 + e, dv[e]);
    __syncwarp();
  }
}

// -------------------------------------------------------------------------------------
// optimizer sweep over EVERY element of EVERY core; re-zeroes the gradient scratch.
// SGD: tt_embeddings_cuda.cu:392; Adagrad: :412-414 (state += g*g; w -= lr*g/(sqrt(state)+eps))
// -------------------------------------------------------------------------------------
struct SweepArgs {
  float* w[TTB_MAX_CORES];
  float* g[TTB_MAX_CORES];
  float* s[TTB_MAX_CORES];
  long long numel[TTB_MAX_CORES];  // multiples of 1 (tail handled scalar)
  int T;
};

// kSweepUnroll independent 16-byte gradient loads are in flight per thread before the first one is looked at (the
// sweep is a chain of dependent loads -- gradient, then weight, then state -- over a few MB: latency, not bandwidth,
// is what a small table pays, and a large one needs the bytes in flight to fill HBM).
constexpr int kSweepUnroll = 4;

template <bool ADAGRAD>
__global__ void __launch_bounds__(256)
    optimizer_sweep_kernel(const SweepArgs a, const float lr, const float eps, const int pdl) {
  pdl_wait(pdl);  // every gradient contribution of the backward kernel has landed
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (int t = 0; t < a.T; ++t) {
    float* __restrict__ w = a.w[t];
    float* __restrict__ g = a.g[t];
    float* __restrict__ s = a.s[t];
    const long long n = a.numel[t];
    const long long n4 = n >> 2;
    for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x; base < n4; base += stride * kSweepUnroll) {
      float4 gv[kSweepUnroll], wv[kSweepUnroll], sv[kSweepUnroll];
      bool live[kSweepUnroll];
#pragma unroll
      for (int u = 0; u < kSweepUnroll; ++u) {
        const long long i = base + u * stride;
        gv[u] = i < n4 ? reinterpret_cast<float4*>(g)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < kSweepUnroll; ++u) {
        const long long i = base + u * stride;
        live[u] = !(gv[u].x == 0.f && gv[u].y == 0.f && gv[u].z == 0.f && gv[u].w == 0.f);  // untouched: skip
        if (live[u]) {
          wv[u] = reinterpret_cast<float4*>(w)[i];
          if (ADAGRAD) sv[u] = reinterpret_cast<float4*>(s)[i];
        }
      }
#pragma unroll
      for (int u = 0; u < kSweepUnroll; ++u) {
        if (!live[u]) continue;
        const long long i = base + u * stride;
        if (ADAGRAD) {
          sv[u].x += gv[u].x * gv[u].x;
          sv[u].y += gv[u].y * gv[u].y;
          sv[u].z += gv[u].z * gv[u].z;
          sv[u].w += gv[u].w * gv[u].w;
          wv[u].x -= lr * gv[u].x / (sqrtf(sv[u].x) + eps);
          wv[u].y -= lr * gv[u].y / (sqrtf(sv[u].y) + eps);
          wv[u].z -= lr * gv[u].z / (sqrtf(sv[u].z) + eps);
          wv[u].w -= lr * gv[u].w / (sqrtf(sv[u].w) + eps);
          reinterpret_cast<float4*>(s)[i] = sv[u];
        } else {
          wv[u].x -= lr * gv[u].x;
          wv[u].y -= lr * gv[u].y;
          wv[u].z -= lr * gv[u].z;
          wv[u].w -= lr * gv[u].w;
        }
        reinterpret_cast<float4*>(w)[i] = wv[u];
        reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    // scalar tail (numel % 4)
    const long long tail0 = n4 << 2;
    for (long long i = tail0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
      const float gv = g[i];
      if (gv == 0.f) continue;
      if (ADAGRAD) {
        const float sv = s[i] + gv * gv;
        s[i] = sv;
        w[i] -= lr * gv / (sqrtf(sv) + eps);
      } else {
        w[i] -= lr * gv;
      }
      g[i] = 0.f;
    }
  }
}

}  // namespace

int launch_fwd_generic(const ChainDims& d, int64_t nnz, const int64_t* indices,
                       const int64_t* rowidx, const int64_t* tableidx, const CorePtrs& cores,
                       float* output, const int32_t* mask, cudaStream_t stream) {
  const size_t smem = (size_t)kFwdWarps * 2 * d.vmax * sizeof(float);
  TTB_CHECK(smem <= 227 * 1024, "tt_forward(generic): chain state of %zu B exceeds shared memory",
            smem);
  static SmemAttr attr;
  TTB_CUDA(attr.ensure(tt_fwd_generic_kernel, smem));
  long long blocks = (nnz + kFwdWarps - 1) / kFwdWarps;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  KernelTimer timer(TTB_KIND_FWD, stream);
  tt_fwd_generic_kernel<<<(unsigned)blocks, kFwdWarps * kWarp, smem, stream>>>(
      d, nnz, (const long long*)indices, (const long long*)rowidx, (const long long*)tableidx,
      cores, output, mask);
  TTB_LAUNCH_CHECK();
  return 0;
}

int launch_bwd_generic(const ChainDims& d, int64_t nnz, const int64_t* indices,
                       const int64_t* rowidx, const int64_t* tableidx, const float* d_output,
                       const CorePtrs& cores, const CorePtrsRW& grads, const int32_t* mask,
                       cudaStream_t stream) {
  const size_t smem = (size_t)kBwdWarps * (d.vsum + 2 * d.vmax) * sizeof(float);
  TTB_CHECK(smem <= 227 * 1024, "tt_backward(generic): chain state of %zu B exceeds shared memory",
            smem);
  static SmemAttr attr;
  TTB_CUDA(attr.ensure(tt_bwd_generic_kernel, smem));
  long long blocks = (nnz + kBwdWarps - 1) / kBwdWarps;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  KernelTimer timer(TTB_KIND_BWD, stream);
  tt_bwd_generic_kernel<<<(unsigned)blocks, kBwdWarps * kWarp, smem, stream>>>(
      d, nnz, (const long long*)indices, (const long long*)rowidx, (const long long*)tableidx,
      d_output, cores, grads, mask);
  TTB_LAUNCH_CHECK();
  return 0;
}

int launch_optimizer_sweep(const ChainDims& d, int optim, float lr, float eps,
                           const CorePtrsRW& cores, const CorePtrsRW& grads,
                           const CorePtrsRW& state, cudaStream_t stream, int core_mask) {
  SweepArgs a;
  a.T = d.T;
  long long total = 0;
  for (int t = 0; t < TTB_MAX_CORES; ++t) {
    a.w[t] = t < d.T ? cores.c[t] : nullptr;
    a.g[t] = t < d.T ? grads.c[t] : nullptr;
    a.s[t] = (t < d.T && optim == TTB_OPTIM_ADAGRAD) ? state.c[t] : nullptr;
    a.numel[t] = (t < d.T && ((core_mask >> t) & 1)) ? (long long)d.num_tables * d.p[t] * d.S[t] : 0;
    total += a.numel[t];
  }
  long long blocks = (total / 4 + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  KernelTimer timer(TTB_KIND_SWEEP, stream);
  const bool pdl = tuning_flag("TTB_PDL");
  if (optim == TTB_OPTIM_ADAGRAD)
    TTB_CUDA(launch_kernel(pdl, optimizer_sweep_kernel<true>, dim3((unsigned)blocks), dim3(256), 0, stream, a, lr,
                           eps, pdl ? 1 : 0));
  else
    TTB_CUDA(launch_kernel(pdl, optimizer_sweep_kernel<false>, dim3((unsigned)blocks), dim3(256), 0, stream, a, lr,
                           eps, pdl ? 1 : 0));
  TTB_LAUNCH_CHECK();
  return 0;
}

}  // namespace ttb
