// Round-2 probe (written without a GPU at hand; run it first thing on a B200):
//   nvcc -gencode arch=compute_100a,code=sm_100a -I fbtt_embedding_b200/csrc -I include -o /tmp/mma_probe2 tests/cuda/mma_probe2.cu
// Two tcgen05 features the backward kernel's next revision depends on:
//
//  (1) MN-major tf32 operands.  tests/cuda/mma_probe.cu found that MN-major tf32 with the standard 128B swizzle
//      (layout type 2) yields all-zero accumulators.  CuTe's Layout_MN_SW128_32B_Atom says the canonical MN-major
//      layout for 32-bit types is layout type 1, SWIZZLE_128B_BASE32B: 128-byte rows (32 elements of MN), swizzle
//      atom = 4 K-rows, the 32-BYTE chunk index of a row XOR-ed with (k & 3)  (Swizzle<2,5,2> on byte addresses);
//      LBO = stride between 32-wide MN blocks, SBO = stride between 4-row K groups, one K=8 MMA step spans two
//      K groups.  If it works, B1 / G / A0 can be staged in their natural row-major order for the GEMMs that
//      need them "transposed" (no scalar transposing stores).
//  (2) A operand from TMEM (tcgen05.mma [d], [a_tmem], b_desc, ...).  G = dOut * C2^T is produced by the SIMT
//      stage one tile row per thread, i.e. already in TMEM-lane order: tcgen05.st it and feed MMA-3 from TMEM,
//      which frees the 64 KB K-major copy of G in shared memory (room to double-buffer the gather stage).
//
// Every variant computes D[128 x N] = A[128 x K] * B[K x N] on small integers (exact in tf32) and prints the
// max abs error against the host; a variant that traps or hangs is bounded by mbar_wait's spin limit.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ttb_sm100.cuh"
using namespace ttb::sm100;

struct Variant {
  int mode;      // 0: B MN-major BASE32B   1: A MN-major BASE32B   2: A from TMEM (B K-major SW128)
  int N, K;      // N multiple of 32 (<= 128), K multiple of 8 (<= 128)
  int swap;      // swap the LBO / SBO roles
  int sbo;       // bytes between 4-row K groups (dense: 512)
};

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

__device__ __forceinline__ void mma_tf32_a_tmem(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
               "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
               "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// a_img / b_img: shared-memory images prepared on the host; a_rowmajor: A[128][K] plain (mode 2)
__global__ void probe2_kernel(const float* a_img, int a_bytes, const float* b_img, int b_bytes, Variant v, float* d_out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + ((a_bytes + 1023) & ~1023);
  __shared__ uint64_t mbar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = 128;
  if (v.mode != 2)
    for (int i = tid; i < a_bytes / 4; i += blockDim.x) ((float*)sA)[i] = a_img[i];
  for (int i = tid; i < b_bytes / 4; i += blockDim.x) ((float*)sB)[i] = b_img[i];
  if (warp == 0) tmem_alloc<512>(&slot);
  if (tid == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
  fence_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tD = slot, tA = slot + 256;
  if (v.mode == 2) {  // A[row][0..K) -> TMEM lane row, columns tA .. tA+K
    const int row = warp * 32 + lane;
    for (int c = 0; c < v.K; c += 8) {
      float x[8];
      for (int i = 0; i < 8; ++i) x[i] = a_img[row * v.K + c + i];
      tmem_st8(tA + ((uint32_t)(warp * 32) << 16) + c, x);
    }
    tmem_st_wait();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
  }
  const uint32_t idesc = make_idesc_tf32(M, v.N, v.mode == 1, v.mode == 0);
  if (tid == 0) {
    for (int ks = 0; ks < v.K / 8; ++ks) {
      // K-major SW128 operand stored [MN rows][K cols]: 32-column blocks of rows*128 B, +32 B per K step inside
      const uint64_t a_k = make_desc(smem_u32(sA) + (ks / 4) * (M * 128) + (ks % 4) * 32, 16, 1024, 2);
      const uint64_t b_k = make_desc(smem_u32(sB) + (ks / 4) * (v.N * 128) + (ks % 4) * 32, 16, 1024, 2);
      // MN-major BASE32B operand stored [K rows][MN cols]: one K step = 8 rows = two 4-row swizzle groups
      uint32_t lbo = (uint32_t)v.K * 128, sbo = (uint32_t)v.sbo;
      if (v.swap) { const uint32_t t = lbo; lbo = sbo; sbo = t; }
      const uint64_t a_mn = make_desc(smem_u32(sA) + ks * 1024, lbo, sbo, 1);
      const uint64_t b_mn = make_desc(smem_u32(sB) + ks * 1024, lbo, sbo, 1);
      if (v.mode == 0) mma_tf32(tD, a_k, b_mn, idesc, ks > 0);
      if (v.mode == 1) mma_tf32(tD, a_mn, b_k, idesc, ks > 0);
      if (v.mode == 2) mma_tf32_a_tmem(tD, tA + ks * 8, b_k, idesc, ks > 0);
    }
    mma_commit(&mbar);
  }
  mbar_wait(&mbar, 0);
  tc_fence_after_sync();
  if (warp < 4) {
    for (int c = 0; c < v.N; c += 16) {
      float r[16];
      tmem_ld16(tD + ((uint32_t)(warp * 32) << 16) + c, r);
      tmem_ld_wait();
      for (int i = 0; i < 16; ++i) d_out[(warp * 32 + lane) * v.N + c + i] = r[i];
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(slot);
}

// K-major SW128: X[rows][cols], 16-byte chunk XOR (row & 7)
static size_t off_k_sw128(int rows, int r, int col) {
  const int b = col >> 5, c = (col >> 2) & 7;
  return (size_t)b * rows * 128 + (size_t)r * 128 + ((c ^ (r & 7)) << 4) + ((col & 3) << 2);
}
// MN-major SW128_BASE32B: X[K rows][MN cols], 32-byte chunk XOR (row & 3); blocks of 32 MN columns
static size_t off_mn_b32(int krows, int k, int mn) {
  const int b = mn >> 5, c = (mn >> 3) & 3;
  return (size_t)b * krows * 128 + (size_t)k * 128 + ((c ^ (k & 3)) << 5) + ((mn & 7) << 2);
}

int main() {
  const int M = 128;
  std::vector<Variant> vs = {
      {0, 128, 32, 0, 512}, {0, 128, 32, 1, 512}, {0, 32, 128, 0, 512}, {0, 32, 128, 1, 512},  // B MN-major
      {1, 32, 128, 0, 512}, {1, 32, 128, 1, 512}, {1, 128, 32, 0, 512},                       // A MN-major
      {2, 32, 128, 0, 512}, {2, 128, 32, 0, 512},                                             // A from TMEM
  };
  int failures = 0;
  for (auto& v : vs) {
    std::vector<float> A(M * v.K), B(v.K * v.N), D(M * v.N, 0.f), Dref(M * v.N, 0.f);
    for (int i = 0; i < M * v.K; ++i) A[i] = (float)((i * 7 + 3) % 11 - 5);
    for (int i = 0; i < v.K * v.N; ++i) B[i] = (float)((i * 5 + 1) % 13 - 6);
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < v.N; ++n) {
        float s = 0;
        for (int k = 0; k < v.K; ++k) s += A[m * v.K + k] * B[k * v.N + n];
        Dref[m * v.N + n] = s;
      }
    std::vector<float> ai((size_t)M * v.K, 0.f), bi((size_t)v.K * v.N, 0.f);
    for (int m = 0; m < M; ++m)
      for (int k = 0; k < v.K; ++k) {
        if (v.mode == 0) ai[off_k_sw128(M, m, k) / 4] = A[m * v.K + k];          // K-major A [M][K]
        else if (v.mode == 1) ai[off_mn_b32(v.K, k, m) / 4] = A[m * v.K + k];   // MN-major A stored [K][M]
        else ai[(size_t)m * v.K + k] = A[m * v.K + k];                          // plain, goes to TMEM
      }
    for (int k = 0; k < v.K; ++k)
      for (int n = 0; n < v.N; ++n) {
        if (v.mode == 0) bi[off_mn_b32(v.K, k, n) / 4] = B[k * v.N + n];        // MN-major B stored [K][N]
        else bi[off_k_sw128(v.N, n, k) / 4] = B[k * v.N + n];                   // K-major B stored [N][K]
      }
    float *da, *db, *dd;
    cudaMalloc(&da, ai.size() * 4); cudaMalloc(&db, bi.size() * 4); cudaMalloc(&dd, D.size() * 4);
    cudaMemcpy(da, ai.data(), ai.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, bi.data(), bi.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0, D.size() * 4);
    const int smem = 1024 + (((int)ai.size() * 4 + 1023) & ~1023) + (int)bi.size() * 4 + 1024;
    cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe2_kernel<<<1, 128, smem>>>(da, (int)ai.size() * 4, db, (int)bi.size() * 4, v, dd);
    const cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0; int nz = 0;
    for (size_t i = 0; i < D.size(); ++i) {
      maxerr = fmax(maxerr, fabs(D[i] - Dref[i])); maxref = fmax(maxref, fabs(Dref[i])); nz += D[i] != 0.f;
    }
    const char* what = v.mode == 0 ? "B mn-major base32b" : v.mode == 1 ? "A mn-major base32b" : "A from TMEM";
    printf("%-20s N=%3d K=%3d swap=%d sbo=%d : %s maxerr=%g (max|ref|=%g) nonzero=%d/%zu %s\n", what, v.N, v.K, v.swap,
           v.sbo, cudaGetErrorString(e), maxerr, maxref, nz, D.size(), maxerr == 0 ? "OK" : "MISMATCH");
    failures += maxerr != 0;
    cudaFree(da); cudaFree(db); cudaFree(dd);
    if (e != cudaSuccess) { printf("aborting after CUDA error\n"); return 2; }
  }
  return failures ? 1 : 0;
}
