// Round-2 probe (written without a GPU at hand; run it with the first GPU call of round 2):
//   nvcc -gencode arch=compute_100a,code=sm_100a -I fbtt_embedding_b200/csrc -I include -o /tmp/mma_probe3 tests/cuda/mma_probe3.cu
// bf16 operands for the tcgen05 kernels (tcgen05.mma kind::f16, a/b format = BF16, fp32 accumulate in TMEM) -- the
// route to ranks >= 64 on tcgen05 and to a backward kernel that fits two CTAs per SM:
//
//  (1) layouts.  16-bit operands have BOTH majors under the standard 128B swizzle (layout type 2), unlike tf32
//      (tests/cuda/mma_probe.cu, mma_probe2.cu): one staged tile X[rows][64 bf16 = 128 B] serves as the K-major
//      operand [MN = rows][K = cols] and as the MN-major operand [K = rows][MN = cols] of the transposed GEMM, so
//      the backward needs no second (transposed) copy of A0 / B1 / G.  Canonical MN-major SW128 layout (CuTe,
//      mma_traits_sm100.hpp): ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units -- 128-byte rows of 64 MN elements,
//      16-byte chunk XOR (k & 7), 8-row K groups SBO apart, 64-wide MN blocks LBO apart; one K = 16 MMA step spans
//      two K groups.
//  (2) precision.  bf16 keeps 8 mantissa bits, tf32 11: a single bf16 product misses the 1e-3 forward tolerance.
//      Splitting x = hi + lo (hi = bf16(x), lo = bf16(x - hi)) and accumulating Ahi*Bhi + Ahi*Blo + Alo*Bhi in the
//      same TMEM tile keeps ~16 bits at three times the (idle) tensor work.  The probe prints the max-norm relative
//      error of 1 / 3 / 4 MMAs against an fp64 reference, next to tf32's.
//
// Exact cases use small integers (exact in bf16); a variant that traps or hangs is bounded by mbar_wait's spin limit.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "ttb_sm100.cuh"
using namespace ttb::sm100;

struct Variant {
  int a_mn, b_mn;  // 1: operand staged MN-major ([K rows][MN cols]), 0: K-major ([MN rows][K cols])
  int N, K;        // N multiple of 64 (<= 128), K multiple of 64 (<= 128)
  int swap;        // swap the LBO / SBO roles of the MN-major descriptors
  int terms;       // 1: hi*hi   3: + hi*lo + lo*hi   4: + lo*lo   (split-precision accumulation)
  int tf32;        // 1: kind::tf32 single MMA on the same data (K-major only), for the error comparison
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
// instruction descriptor, kind::f16: c_format [4,6) = 1 (F32); a_format [7,10) = 1 (BF16); b_format [10,13) = 1;
// a_major bit 15, b_major bit 16; N >> 3 in [17,23); M >> 4 in [24,29)   (cute/arch/mma_sm100_desc.hpp)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn, int b_mn) {
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

// images: [A hi | A lo | B hi | B lo], each img_bytes (bf16) -- or fp32 images of twice the size for the tf32 run
__global__ void probe3_kernel(const uint8_t* img, int a_bytes, int b_bytes, Variant v, float* d_out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  const int a_al = (a_bytes + 1023) & ~1023, b_al = (b_bytes + 1023) & ~1023;
  const int halves = v.tf32 ? 1 : 2;  // the tf32 run stages one fp32 image per operand (twice the bytes of a bf16 one)
  uint8_t* sA[2] = {smem, smem + (halves - 1) * a_al};
  uint8_t* sB[2] = {smem + halves * a_al, smem + halves * a_al + (halves - 1) * b_al};
  __shared__ uint64_t mbar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = 128;
  for (int h = 0; h < halves; ++h) {
    for (int i = tid; i < a_bytes / 4; i += blockDim.x) ((uint32_t*)sA[h])[i] = ((const uint32_t*)(img + (size_t)h * a_bytes))[i];
    for (int i = tid; i < b_bytes / 4; i += blockDim.x)
      ((uint32_t*)sB[h])[i] = ((const uint32_t*)(img + (size_t)halves * a_bytes + (size_t)h * b_bytes))[i];
  }
  if (warp == 0) tmem_alloc<128>(&slot);
  if (tid == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
  fence_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tD = slot;
  if (tid == 0) {
    if (v.tf32) {  // fp32 images, K-major SW128, 32-element (128 B) K blocks, K = 8 per MMA
      const uint32_t idesc = make_idesc_tf32(M, v.N, 0, 0);
      for (int ks = 0; ks < v.K / 8; ++ks) {
        const uint64_t a = make_desc(smem_u32(sA[0]) + (ks / 4) * (M * 128) + (ks % 4) * 32, 16, 1024, 2);
        const uint64_t b = make_desc(smem_u32(sB[0]) + (ks / 4) * (v.N * 128) + (ks % 4) * 32, 16, 1024, 2);
        mma_tf32(tD, a, b, idesc, ks > 0);
      }
    } else {
      const uint32_t idesc = make_idesc_bf16(M, v.N, v.a_mn, v.b_mn);
      // split-precision terms in order of decreasing magnitude: (hi,hi) (hi,lo) (lo,hi) (lo,lo)
      const int ta[4] = {0, 0, 1, 1}, tb[4] = {0, 1, 0, 1};
      bool first = true;
      for (int term = 0; term < v.terms; ++term) {
        for (int ks = 0; ks < v.K / 16; ++ks) {
          // K-major [MN rows][K cols]: 64-element (128 B) K blocks of rows*128 B, K = 16 step = +32 B inside a block
          const uint64_t a_k = make_desc(smem_u32(sA[ta[term]]) + (ks / 4) * (M * 128) + (ks % 4) * 32, 16, 1024, 2);
          const uint64_t b_k = make_desc(smem_u32(sB[tb[term]]) + (ks / 4) * (v.N * 128) + (ks % 4) * 32, 16, 1024, 2);
          // MN-major [K rows][MN cols]: 64-element MN blocks of K*128 B (LBO), 8-row K groups 1024 B apart (SBO);
          // a K = 16 step starts 16 rows = 2048 B further down
          uint32_t lbo = (uint32_t)v.K * 128, sbo = 1024;
          if (v.swap) { const uint32_t t = lbo; lbo = sbo; sbo = t; }
          const uint64_t a_mn = make_desc(smem_u32(sA[ta[term]]) + ks * 2048, lbo, sbo, 2);
          const uint64_t b_mn = make_desc(smem_u32(sB[tb[term]]) + ks * 2048, lbo, sbo, 2);
          mma_bf16(tD, v.a_mn ? a_mn : a_k, v.b_mn ? b_mn : b_k, idesc, first ? 0u : 1u);
          first = false;
        }
      }
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
  if (warp == 0) tmem_dealloc<128>(slot);
}

// ---- host ----------------------------------------------------------------------------------------------------
static uint16_t bf16_rn(float x) {  // round to nearest even
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static float bf16_to_f(uint16_t h) {
  const uint32_t u = (uint32_t)h << 16;
  float x;
  memcpy(&x, &u, 4);
  return x;
}
// X[rows][cols] with 128-byte rows per 64-column (bf16) / 32-column (fp32) block, 16-byte chunk XOR (row & 7):
// the K-major image of [MN = rows][K = cols] AND the MN-major image of [K = rows][MN = cols]
static size_t off_sw128(int rows, int r, int col, int elem_bytes) {
  const int per_row = 128 / elem_bytes, per_chunk = 16 / elem_bytes;
  const int b = col / per_row, c = (col % per_row) / per_chunk;
  return (size_t)b * rows * 128 + (size_t)r * 128 + (size_t)((c ^ (r & 7)) << 4) + (size_t)(col % per_chunk) * elem_bytes;
}

int main() {
  const int M = 128;
  std::vector<Variant> vs = {
      // exact integer cases: layouts
      {0, 0, 128, 64, 0, 1, 0}, {0, 0, 64, 128, 0, 1, 0},                            // both K-major (sanity)
      {0, 1, 128, 64, 0, 1, 0}, {0, 1, 128, 64, 1, 1, 0}, {0, 1, 64, 128, 0, 1, 0},  // B MN-major
      {1, 0, 64, 128, 0, 1, 0}, {1, 0, 64, 128, 1, 1, 0}, {1, 0, 128, 64, 0, 1, 0},  // A MN-major
      {1, 1, 128, 128, 0, 1, 0},                                                     // both MN-major
      // random fp32 data: precision of 1 / 3 / 4 bf16 terms and of tf32
      {0, 0, 128, 128, 0, 1, 0}, {0, 0, 128, 128, 0, 3, 0}, {0, 0, 128, 128, 0, 4, 0}, {0, 0, 128, 128, 0, 1, 1},
  };
  int failures = 0, idx = 0;
  for (auto& v : vs) {
    const bool exact = idx++ < 9;
    std::vector<float> A(M * v.K), B(v.K * v.N), D(M * v.N, 0.f);
    std::vector<double> Dref(M * v.N, 0.0);
    srand(7);
    for (int i = 0; i < M * v.K; ++i)
      A[i] = exact ? (float)((i * 7 + 3) % 11 - 5) : ((float)rand() / RAND_MAX - 0.5f);
    for (int i = 0; i < v.K * v.N; ++i)
      B[i] = exact ? (float)((i * 5 + 1) % 13 - 6) : ((float)rand() / RAND_MAX - 0.5f);
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < v.N; ++n) {
        double s = 0;
        for (int k = 0; k < v.K; ++k) s += (double)A[m * v.K + k] * (double)B[k * v.N + n];
        Dref[m * v.N + n] = s;
      }
    const int eb = v.tf32 ? 4 : 2;
    const size_t a_bytes = (size_t)M * v.K * eb, b_bytes = (size_t)v.K * v.N * eb;
    const int halves = v.tf32 ? 1 : 2;
    std::vector<uint8_t> img(halves * (a_bytes + b_bytes), 0);
    auto put = [&](size_t base, size_t off, float x, int half) {
      if (v.tf32) {  // round to nearest tf32 as the production kernels do when staging (to_tf32, ttb_sm100.cuh)
        uint32_t u;
        memcpy(&u, &x, 4);
        u = (u + 0x1000u) & 0xffffe000u;
        if (half == 0) memcpy(&img[base + off], &u, 4);
        return;
      }
      const uint16_t hi = bf16_rn(x);
      const uint16_t lo = bf16_rn(x - bf16_to_f(hi));
      const uint16_t w = half ? lo : hi;
      memcpy(&img[base + off], &w, 2);
    };
    for (int half = 0; half < halves; ++half) {
      for (int m = 0; m < M; ++m)
        for (int k = 0; k < v.K; ++k)
          put((size_t)half * a_bytes, v.a_mn ? off_sw128(v.K, k, m, eb) : off_sw128(M, m, k, eb), A[m * v.K + k], half);
      for (int k = 0; k < v.K; ++k)
        for (int n = 0; n < v.N; ++n)
          put((size_t)halves * a_bytes + (size_t)half * b_bytes, v.b_mn ? off_sw128(v.K, k, n, eb) : off_sw128(v.N, n, k, eb),
              B[k * v.N + n], half);
    }
    uint8_t* dimg;
    float* dd;
    cudaMalloc(&dimg, img.size());
    cudaMalloc(&dd, D.size() * 4);
    cudaMemcpy(dimg, img.data(), img.size(), cudaMemcpyHostToDevice);
    cudaMemset(dd, 0, D.size() * 4);
    const int smem = 1024 + halves * ((int)((a_bytes + 1023) & ~(size_t)1023) + (int)((b_bytes + 1023) & ~(size_t)1023)) + 1024;
    cudaFuncSetAttribute(probe3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe3_kernel<<<1, 128, smem>>>(dimg, (int)a_bytes, (int)b_bytes, v, dd);
    const cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    int nz = 0;
    for (size_t i = 0; i < D.size(); ++i) {
      maxerr = fmax(maxerr, fabs((double)D[i] - Dref[i]));
      maxref = fmax(maxref, fabs(Dref[i]));
      nz += D[i] != 0.f;
    }
    if (exact) {
      printf("bf16 A %s  B %s  N=%3d K=%3d swap=%d : %s maxerr=%g (max|ref|=%g) nonzero=%d/%zu %s\n",
             v.a_mn ? "MN-major" : "K-major ", v.b_mn ? "MN-major" : "K-major ", v.N, v.K, v.swap, cudaGetErrorString(e),
             maxerr, maxref, nz, D.size(), maxerr == 0 ? "OK" : "MISMATCH");
      if (!v.swap) failures += maxerr != 0;  // the swapped variants only tell which field is which
    } else {
      printf("%s  N=%3d K=%3d : %s max-norm relative error %.3g (smem %d B)\n",
             v.tf32 ? "tf32, 1 MMA pass       " : v.terms == 1 ? "bf16 hi*hi             "
                                              : v.terms == 3 ? "bf16 hi*hi+hi*lo+lo*hi "
                                                             : "bf16 all four terms    ",
             v.N, v.K, cudaGetErrorString(e), maxerr / maxref, smem);
    }
    cudaFree(dimg);
    cudaFree(dd);
    if (e != cudaSuccess) {
      printf("aborting after CUDA error\n");
      return 2;
    }
  }
  return failures ? 1 : 0;
}
