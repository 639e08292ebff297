// Experiment (not part of the library): tcgen05.mma with an MN-major SWIZZLE_128B B operand whose rows are the K index.
// Image: row = k (a pixel), 128-byte row = 64 consecutive n (channels), 8-row groups of 1024 B, the usual address swizzle;
// a second 64-channel block sits `lbo` bytes further.  D[m][n] = sum_k A[m][k] * B[k + r0][n], A K-major.  Checks
//   (1) b_major = 1 (idesc bit 16) with SBO = 1024 and LBO = block stride, N = 64 and N = 128,
//   (2) K advance by 16 rows (2048 B) between MMAs, (3) an arbitrary ROW offset r0 of the start address (tap shift).
// This is the operand form a haloed weight-gradient kernel needs (contraction over pixels, taps by row offset).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I3dhumangan_b200/csrc tools/experiments/desc_mn_major.cu -o tools/experiments/bin/desc_mn_major
#include <cstdio>
#include <vector>
#include "umma.cuh"

using namespace hg;

constexpr int kRows = 208;                 // K rows in the B image
constexpr uint32_t kImg = 26 * 1024;       // 208 * 128 B

__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ Bm, int r0, int N, int lbo_mode,
                                             float* __restrict__ out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_t = smem;                       // [128 x 64] K-major
  uint8_t* b_t = smem + 16 * 1024;           // two images [kRows x 64]
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
    const int m = i / 64, k = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(a_t + sw128_offset(m, k)) = __float2bfloat16(A[i]);
  }
  for (int i = threadIdx.x; i < kRows * 128; i += blockDim.x) {
    const int k = i / 128, n = i % 128;
    *reinterpret_cast<__nv_bfloat16*>(b_t + (n / 64) * kImg + sw128_offset(k, n % 64)) = __float2bfloat16(Bm[i]);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc<128>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N) | (1u << 16);          // B is MN-major
    const uint64_t da = umma_desc_sw128(smem_u32(a_t));
    uint64_t db = umma_desc_sw128(smem_u32(b_t) + r0 * 128);
    const uint64_t lbo = lbo_mode == 0 ? (kImg >> 4) : 1;
    db = (db & ~(static_cast<uint64_t>(0x3FFF) << 16)) | (lbo << 16);
    for (uint32_t k = 0; k < 4; ++k) umma_bf16(tmem, da + 2 * k, db + 128 * k, idesc, k > 0);     // +16 rows = 2048 B
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 128 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<128>(tmem);
}

int main() {
  std::vector<float> A(128 * 64), Bm(kRows * 128);
  for (int i = 0; i < 128 * 64; ++i) A[i] = static_cast<float>((i * 7) % 5) - 2.f;
  for (int i = 0; i < kRows * 128; ++i) Bm[i] = static_cast<float>((i * 13 + i / 128) % 7) - 3.f;
  float *dA, *dB, *dO;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dB, Bm.size() * 4);
  cudaMalloc(&dO, 128 * 128 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bm.data(), Bm.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  std::vector<float> O(128 * 128);
  for (int N : {64, 128})
    for (int lbo_mode : {0, 1})
      for (int r0 : {0, 1, 7, 8, 129, 131}) {
        cudaMemset(dO, 0, O.size() * 4);
        probe<<<1, 128, 80 * 1024>>>(dA, dB, r0, N, lbo_mode, dO);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N %d lbo %d r0 %d: CUDA error %s\n", N, lbo_mode, r0, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0, first = -1;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < N; ++n) {
            float ref = 0.f;
            for (int k = 0; k < 64; ++k) ref += A[m * 64 + k] * Bm[(k + r0) * 128 + n];
            if (O[m * 128 + n] != ref) { if (first < 0) first = m * 128 + n; ++bad; }
          }
        printf("N %3d lbo=%s r0 %3d: %s (%d mismatches, first m=%d n=%d)\n", N, lbo_mode == 0 ? "image stride" : "1", r0,
               bad ? "WRONG" : "ok", bad, first / 128, first % 128);
      }
  return 0;
}
