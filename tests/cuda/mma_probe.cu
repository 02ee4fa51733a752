// Standalone probe of the tcgen05 descriptor / layout conventions used by libttb's fast path.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I fbtt_embedding_b200/csrc -I include -o /tmp/mma_probe tests/cuda/mma_probe.cu
// Each variant runs D[128 x N] = A[128 x K] * B[K x N] with small-integer data (exact in tf32)
// and reports the max abs error against a host reference.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ttb_sm100.cuh"
using namespace ttb::sm100;

struct Variant {
  int a_mn, b_mn;      // operand majors
  int N, K;            // K multiple of 8; N multiple of 32
  int swap_lbo_sbo;    // for MN-major operands: swap the roles of LBO/SBO
  int delay;           // spin after the mbarrier wait
  int lbo_k;           // LBO bytes for K-major operands
};

// smem images are prepared on the host (already swizzled) and copied in verbatim
__global__ void probe_kernel(const float* a_img, int a_bytes, const float* b_img, int b_bytes, Variant v,
                             float* d_out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + ((a_bytes + 1023) & ~1023);
  __shared__ uint64_t mbar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < a_bytes / 4; i += blockDim.x) ((float*)sA)[i] = a_img[i];
  for (int i = tid; i < b_bytes / 4; i += blockDim.x) ((float*)sB)[i] = b_img[i];
  if (warp == 0) tmem_alloc<256>(&slot);
  if (tid == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
  fence_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = slot;
  const int M = 128;
  const uint32_t idesc = make_idesc_tf32(M, v.N, v.a_mn, v.b_mn);
  if (tid == 0) {
    for (int ks = 0; ks < v.K / 8; ++ks) {
      uint64_t ad, bd;
      if (!v.a_mn) {
        ad = make_desc_sw128(smem_u32(sA) + (ks / 4) * (M * 128) + (ks % 4) * 32, v.lbo_k, 1024);
      } else {  // stored [K rows][M cols]
        uint32_t lbo = v.K * 128, sbo = 1024;
        if (v.swap_lbo_sbo) { uint32_t t = lbo; lbo = sbo; sbo = t; }
        ad = make_desc_sw128(smem_u32(sA) + ks * 1024, lbo, sbo);
      }
      if (!v.b_mn) {  // stored [N rows][K cols]
        bd = make_desc_sw128(smem_u32(sB) + (ks / 4) * (v.N * 128) + (ks % 4) * 32, v.lbo_k, 1024);
      } else {  // stored [K rows][N cols]
        uint32_t lbo = v.K * 128, sbo = 1024;
        if (v.swap_lbo_sbo) { uint32_t t = lbo; lbo = sbo; sbo = t; }
        bd = make_desc_sw128(smem_u32(sB) + ks * 1024, lbo, sbo);
      }
      mma_tf32(tb, ad, bd, idesc, ks > 0);
    }
    mma_commit(&mbar);
  }
  mbar_wait(&mbar, 0);
  tc_fence_after_sync();
  if (v.delay) __nanosleep(200000);
  if (warp < 4) {
    for (int c = 0; c < v.N; c += 16) {
      float r[16];
      tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c, r);
      tmem_ld_wait();
      for (int i = 0; i < 16; ++i) d_out[(warp * 32 + lane) * v.N + c + i] = r[i];
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tb);
}

static size_t sw_off(int rows, int r, int col) {
  int b = col >> 5, c = (col >> 2) & 7;
  return (size_t)b * rows * 128 + r * 128 + ((c ^ (r & 7)) << 4) + ((col & 3) << 2);
}

int main() {
  std::vector<Variant> vs = {
      {0, 0, 128, 32, 0, 0, 16}, {0, 0, 128, 32, 0, 1, 16}, {0, 0, 128, 32, 0, 0, 0},
      {0, 1, 128, 32, 0, 0, 16}, {0, 1, 128, 32, 1, 0, 16}, {0, 1, 128, 32, 0, 1, 16},
      {0, 0, 32, 128, 0, 0, 16}, {1, 1, 32, 128, 0, 0, 16}, {1, 1, 32, 128, 1, 0, 16},
      {1, 0, 32, 128, 0, 0, 16}, {0, 1, 32, 128, 0, 0, 16},
  };
  const int M = 128;
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
    // smem images
    int a_rows = v.a_mn ? v.K : M, a_cols = v.a_mn ? M : v.K;
    int b_rows = v.b_mn ? v.K : v.N, b_cols = v.b_mn ? v.N : v.K;
    std::vector<float> ai((size_t)a_rows * a_cols, 0.f), bi((size_t)b_rows * b_cols, 0.f);
    for (int r = 0; r < a_rows; ++r)
      for (int c = 0; c < a_cols; ++c) {
        float val = v.a_mn ? A[c * v.K + r] : A[r * v.K + c];
        ai[sw_off(a_rows, r, c) / 4] = val;
      }
    for (int r = 0; r < b_rows; ++r)
      for (int c = 0; c < b_cols; ++c) {
        float val = v.b_mn ? B[r * v.N + c] : B[c * v.N + r];
        bi[sw_off(b_rows, r, c) / 4] = val;
      }
    float *da, *db, *dd;
    cudaMalloc(&da, ai.size() * 4); cudaMalloc(&db, bi.size() * 4); cudaMalloc(&dd, D.size() * 4);
    cudaMemcpy(da, ai.data(), ai.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, bi.data(), bi.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0, D.size() * 4);
    int smem = 1024 + (((int)ai.size() * 4 + 1023) & ~1023) + (int)bi.size() * 4 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<<<1, 128, smem>>>(da, (int)ai.size() * 4, db, (int)bi.size() * 4, v, dd);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0; int nz = 0;
    for (size_t i = 0; i < D.size(); ++i) {
      maxerr = fmax(maxerr, fabs(D[i] - Dref[i])); maxref = fmax(maxref, fabs(Dref[i])); nz += D[i] != 0.f;
    }
    printf("a_mn=%d b_mn=%d N=%d K=%d swap=%d delay=%d lbo_k=%d : %s maxerr=%g (max|ref|=%g) nonzero=%d/%zu  D[0..3]=%g %g %g %g ref=%g %g %g %g\n",
           v.a_mn, v.b_mn, v.N, v.K, v.swap_lbo_sbo, v.delay, v.lbo_k, cudaGetErrorString(e), maxerr, maxref, nz,
           D.size(), D[0], D[1], D[2], D[3], Dref[0], Dref[1], Dref[2], Dref[3]);
    cudaFree(da); cudaFree(db); cudaFree(dd);
    if (e != cudaSuccess) { printf("aborting after CUDA error\n"); return 1; }
  }
  return 0;
}
