// Experiment (not part of the library): does a K-major SWIZZLE_128B UMMA operand descriptor whose start address is offset
// by r0 ROWS (r0 * 128 bytes, not a multiple of the 1024-byte swizzle pattern) address rows r0 .. r0+127 of a tile that was
// written with the swizzle of its ABSOLUTE row index?  If yes, a 3x3 convolution can read all 9 taps from ONE haloed
// operand tile by moving the descriptor start instead of building 9 shifted copies.
// Variants: base_offset field (bits 49-51) = 0, = r0 & 7, = (8 - r0) & 7.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I3dhumangan_b200/csrc tools/experiments/desc_row_offset.cu -o tools/experiments/bin/desc_row_offset
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "umma.cuh"

using namespace hg;

__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, int rows, int r0, int mode, float* __restrict__ out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_t = smem;                       // rows x 64 bf16, SW128 by absolute row
  uint8_t* b_t = smem + 32 * 1024;           // 64 x 64 identity
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < rows * 64; i += blockDim.x) {
    const int r = i / 64, k = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(a_t + sw128_offset(r, k)) = __float2bfloat16(A[i]);
  }
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int n = i / 64, k = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(b_t + sw128_offset(n, k)) = __float2bfloat16(n == k ? 1.f : 0.f);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc<64>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, 64);
    uint64_t da = umma_desc_sw128(smem_u32(a_t) + r0 * 128);
    const uint64_t db = umma_desc_sw128(smem_u32(b_t));
    uint64_t bo = mode == 0 ? 0 : (mode == 1 ? (r0 & 7) : ((8 - r0) & 7));
    da |= bo << 49;
    for (uint32_t k = 0; k < 4; ++k) umma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, k > 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<64>(tmem);
}

int main() {
  const int rows = 160;
  std::vector<float> A(rows * 64);
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < 64; ++k) A[r * 64 + k] = static_cast<float>((r * 64 + k) % 251) - 125.f;   // exact in bf16
  float *dA, *dO;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  std::vector<float> O(128 * 64);
  for (int mode = 0; mode < 3; ++mode)
    for (int r0 : {0, 1, 2, 3, 4, 7, 8, 9, 17, 31}) {
      cudaMemset(dO, 0, O.size() * 4);
      probe<<<1, 128, 48 * 1024>>>(dA, rows, r0, mode, dO);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d r0 %d: CUDA error %s\n", mode, r0, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0, first = -1;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n)
          if (O[m * 64 + n] != A[(m + r0) * 64 + n]) { if (first < 0) first = m * 64 + n; ++bad; }
      printf("mode %d (base_offset %s) r0 %2d: %s (%d mismatches, first at m=%d n=%d)\n", mode,
             mode == 0 ? "0" : (mode == 1 ? "r0&7" : "(8-r0)&7"), r0, bad ? "WRONG" : "ok", bad, first / 64, first % 64);
    }
  return 0;
}
