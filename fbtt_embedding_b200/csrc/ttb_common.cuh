// Shared host/device helpers of libttb (B200 / sm_100a TT-EmbeddingBag).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "ttb.h"

namespace ttb {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int current_path();

// brackets one kernel launch with CUDA events when ttb_timing_enable(1) is active
struct KernelTimer {
  KernelTimer(int kind, cudaStream_t stream);
  ~KernelTimer();
  int slot_;
  cudaStream_t stream_;
};

#define TTB_CHECK(cond, ...)        \
  do {                              \
    if (!(cond)) {                  \
      ::ttb::set_error(__VA_ARGS__); \
      return 1;                     \
    }                               \
  } while (0)

#define TTB_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::ttb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                       __LINE__);                                                        \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

#define TTB_LAUNCH_CHECK()                  \
  do {                                      \
    ::ttb::count_launch();                  \
    TTB_CUDA(cudaPeekAtLastError());        \
  } while (0)

// Opt-in tuning switches read once from the environment (value "1" = on).  Variants that have not been
// measured on a B200 yet stay off by default; scripts/ab_variants.py runs the step under each of them.
bool tuning_flag(const char* name);

constexpr int kWarp = 32;

// Everything a kernel needs to walk the TT chain of one table family, by value.
struct ChainDims {
  int T;
  int num_tables, B, D;
  int p[TTB_MAX_CORES], q[TTB_MAX_CORES], R[TTB_MAX_CORES + 1];
  long long L[TTB_MAX_CORES];
  int S[TTB_MAX_CORES];      // slice elements r_t*q_t*r_{t+1}
  int m[TTB_MAX_CORES];      // rows of v_t: q_0*...*q_t
  int n[TTB_MAX_CORES];      // cols of the link-t GEMM: q_t*r_{t+1}
  int vsize[TTB_MAX_CORES];  // m[t]*R[t+1]
  int vmax;                  // max vsize, rounded up to 4
  int voff[TTB_MAX_CORES];   // prefix offsets of v_0..v_{T-2} (backward keeps them all)
  int vsum;                  // sum of vsize[0..T-2], each rounded up to 4
  long long total_rows;      // prod(p): indices >= this are invalid
  int small32;               // total_rows < 2^31: digit arithmetic may use 32-bit division
  // Heterogeneous table batch (ttb_tt_*_het): tables share q / ranks but have their own p-shapes; their cores are
  // concatenated along the slice dimension, so to every kernel that walks core slices this is ONE table whose
  // p[t] is the concatenated slice count (num_tables == 1 here).  Only the index decomposition knows better:
  // het[tb] (device memory) holds table tb's own p / L / rows and the first slice of the table in each core.
  const ttb_het_table_t* het;
  int het_tables;
  // Fused exchange (ttb_row_map_t): pooled rows go to / gradients come from the batch-slice buffers of the peer
  // ranks instead of [table][row][:] of a local tensor.  Active iff map_peer_off != nullptr.
  const long long* map_peer_off;  // [world] element offset of rank w's buffer relative to the local one
  const int* map_gid;             // [tables] global table number
  int map_bw, map_tt;             // rows per rank, tables of all ranks
};

// element offset of (table, row)'s pooled row relative to `output` / `d_output`.  Host + device: ttb_row_map_offset
// evaluates the same function on host arrays.
__host__ __device__ __forceinline__ long long out_row_offset(const ChainDims& d, long long tb, long long row) {
  if (!d.map_peer_off) return (tb * d.B + row) * d.D;
  const int w = (int)row / d.map_bw;  // row < B < 2^31
  const int r = (int)row - w * d.map_bw;
  return d.map_peer_off[w] + ((long long)r * d.map_tt + d.map_gid[tb]) * d.D;
}

// mixed-radix digits of `idx` in table tb of a heterogeneous batch -> concatenated slice numbers.  Also compiled for
// the host: ttb_het_digits (include/ttb.h) runs this very function on host descriptors, so the decomposition the
// kernels use is pinned against the oracle without a GPU.
__host__ __device__ __forceinline__ bool het_digits(const ChainDims& d, long long tb, long long idx, int (&i)[TTB_MAX_CORES]) {
  if (tb < 0 || tb >= d.het_tables) return false;
  const ttb_het_table_t* h = d.het + tb;
  if (idx < 0 || idx >= h->rows) return false;
#pragma unroll
  for (int t = 0; t < TTB_MAX_CORES; ++t) {
    if (t < d.T) {
      const long long Lt = h->L[t];
      long long qd;
      if (h->rows < (1LL << 31)) {  // a 64-bit division costs ~10x a 32-bit one
        qd = (long long)((unsigned)idx / (unsigned)Lt);
      } else {
        qd = idx / Lt;
      }
      idx -= qd * Lt;
      i[t] = h->off[t] + (int)qd;
    } else {
      i[t] = 0;
    }
  }
  return true;
}

// One batch of lookups as the entry points receive it: COO triples (the reference's op interface) or CSR offsets
// (rowidx == tableidx == nullptr: bag b = table b / B, row b % B; the bucketed path derives them in its plan).
struct LookupBatch {
  int64_t nnz;
  const int64_t* indices;
  const int64_t* rowidx;
  const int64_t* tableidx;
  const int64_t* offsets;
  int64_t num_bags;
  int B;
  const int32_t* mask;  // cache_locations of the async cache front-end (only -1 is a TT lookup), or nullptr
  int zero_output;      // forward: `output` is uninitialised, the library zero-fills it (TTB_BATCH_ZERO_OUTPUT)
  int bf16_cores;       // cores[t] hold bf16 values (TTB_BATCH_BF16_CORES); gradients / optimizer state stay fp32
};

// release/acquire fence at gpu scope (MEMBAR.ALL.GPU): what the threadfence-reduction protocols of this library need;
// __threadfence() compiles to the sequentially-consistent MEMBAR.SC.GPU, which is several times dearer
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// second half of bag_of_guess: the caller has already loaded offsets[b], offsets[b + 1] of the proportional guess b
__device__ __forceinline__ long long bag_of_probe(const long long* __restrict__ offsets, long long num_bags, long long n,
                                                  long long b, long long ob, long long ob1) {
  if (ob <= n && n < ob1) return b;
  long long left, right;  // invariant: offsets[left] <= n < offsets[right]
  if (n < ob) {
    right = b;
    left = b - 1;
    long long step = 1;
    while (left > 0 && __ldg(offsets + left) > n) {
      right = left;
      step <<= 1;
      left = left - step < 0 ? 0 : left - step;
    }
  } else {
    left = b + 1;
    right = left + 1;
    long long step = 1;
    while (right < num_bags && __ldg(offsets + right) <= n) {
      left = right;
      step <<= 1;
      right = right + step > num_bags ? num_bags : right + step;
    }
    if (right > num_bags) right = num_bags;
  }
  while (right - left > 1) {
    const long long mid = (left + right) >> 1;
    if (__ldg(offsets + mid) <= n)
      left = mid;
    else
      right = mid;
  }
  return left;
}

// largest b with offsets[b] <= n, for offsets[0] <= n < offsets[num_bags].  Bags of a batch are roughly equally
// long, so the proportional guess is usually right (ONE round trip: both bounds are loaded together); otherwise
// gallop away from the guess, then bisect -- a plain bisection is log2(num_bags) DEPENDENT loads.
__device__ __forceinline__ long long bag_of_guess(const long long* __restrict__ offsets, long long num_bags,
                                                  long long n, long long nnz) {
  // single precision on purpose: FP64 is a slow pipe here, and a guess that is off by one costs one more probe
  long long b = (long long)((float)n * ((float)num_bags / (float)(nnz > 0 ? nnz : 1)));
  b = b < 0 ? 0 : (b > num_bags - 1 ? num_bags - 1 : b);
  const long long ob = __ldg(offsets + b), ob1 = __ldg(offsets + b + 1);
  return bag_of_probe(offsets, num_bags, n, b, ob, ob1);
}

struct CorePtrs {
  const float* c[TTB_MAX_CORES];
};
struct CorePtrsRW {
  float* c[TTB_MAX_CORES];
};

// validates the shape descriptor and fills ChainDims; returns non-zero + error on failure
int make_chain_dims(const ttb_shape_t* s, ChainDims* d);

inline int sm_count() {
  static int n[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int& v = n[dev & 15];
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
  }
  return v;
}

inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE function attribute: remember, per kernel
// (one static `SmemAttr` per launch site) and per device, the largest value already requested.
struct SmemAttr {
  size_t set[16] = {0};
  template <typename K>
  cudaError_t ensure(K kernel, size_t bytes) {
    const int dev = current_device() & 15;
    if (bytes <= 48 * 1024 || bytes <= set[dev]) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) set[dev] = bytes;
    return e;
  }
};

// Programmatic dependent launch (opt-in, TTB_PDL=1).  A kernel launched with the programmatic-serialization
// attribute may start while its predecessor in the stream is still running; griddepcontrol.wait blocks until
// that predecessor has completed and its writes are visible, so everything before the wait (TMEM allocation,
// barrier init, shared-memory clearing) overlaps the predecessor's tail.  The predecessor opens the door early
// with griddepcontrol.launch_dependents; without it the door opens when it exits (plain serialization).
// `on` is a kernel argument: with 0 neither instruction is executed and the launch is an ordinary one.
__device__ __forceinline__ void pdl_wait(int on) {
  if (on) asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void pdl_trigger(int on) {
  if (on) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// <<<grid, block, smem, stream>>> or, with pdl, the same launch carrying the programmatic-serialization attribute
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                 cudaStream_t stream, Args... args) {
  if (!pdl) {
    kernel<<<grid, block, smem, stream>>>(args...);
    return cudaSuccess;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
// 8-byte vector reduction (sm_90+)
__device__ __forceinline__ void red_add_f32x2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
// 16-byte vector reduction (sm_90+): one L2 atomic transaction for 4 floats
__device__ __forceinline__ void red_add_f32x4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

}  // namespace ttb
