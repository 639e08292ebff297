// Probe: cp.async.bulk.tensor.3d (UTMALDG) of an fp32 box with negative / out-of-range coordinates (zero fill), tensor map
// passed (A) as a __grid_constant__ kernel parameter, (B) through global memory.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O2 -o tools/experiments/bin/tma_probe tools/experiments/tma_probe.cu
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

constexpr int BW = 40, BH = 39;

__device__ __forceinline__ void load_box(const void* map, float* dst, uint64_t* bar, int c0, int c1, int c2) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
  const uint32_t mb = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(BW * BH * 4) : "memory");
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(d), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(mb) : "memory");
}

__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  const uint32_t mb = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(mb), "r"(parity) : "memory");
  }
}

template <bool kParam>
__global__ void probe(const __grid_constant__ CUtensorMap pmap, const CUtensorMap* gmap, float* out, int c0, int c1, int c2) {
  __shared__ __align__(128) float tile[BH * BW];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    const uint32_t mb = static_cast<uint32_t>(__cvta_generic_to_shared(&bar));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) load_box(kParam ? static_cast<const void*>(&pmap) : static_cast<const void*>(gmap), tile, &bar, c0, c1, c2);
  wait_bar(&bar, 0);
  for (int i = threadIdx.x; i < BH * BW; i += blockDim.x) out[i] = tile[i];
}

int main() {
  const int P = 4, H = 150, W = 204;
  std::vector<float> h(static_cast<size_t>(P) * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = static_cast<float>(i % 9973) * 0.25f + 1.f;
  float *x, *out;
  cudaMalloc(&x, h.size() * 4);
  cudaMalloc(&out, BH * BW * 4);
  cudaMemcpy(x, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
    printf("no entry point\n");
    return 1;
  }
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  CUtensorMap map;
  const cuuint64_t dims[3] = {W, H, P};
  const cuuint64_t strides[2] = {W * 4ull, static_cast<cuuint64_t>(W) * H * 4};
  const cuuint32_t box[3] = {BW, BH, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", static_cast<int>(r));
  CUtensorMap* gmap;
  cudaMalloc(&gmap, sizeof(map));
  cudaMemcpy(gmap, &map, sizeof(map), cudaMemcpyHostToDevice);
  const int coords[4][3] = {{0, 0, 0}, {-4, -2, 1}, {180, 130, 3}, {-3, -2, 1}};   // the last one: c0 * 4 B not 16-byte aligned
  for (int variant = 0; variant < 2; ++variant)
    for (auto& c : coords) {
      cudaMemset(out, 0xff, BH * BW * 4);
      if (variant == 0) probe<true><<<1, 128>>>(map, gmap, out, c[0], c[1], c[2]);
      else probe<false><<<1, 128>>>(map, gmap, out, c[0], c[1], c[2]);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("variant %s coords (%d,%d,%d): %s\n", variant == 0 ? "param" : "global", c[0], c[1], c[2], cudaGetErrorString(e));
        return 2;
      }
      std::vector<float> o(BH * BW);
      cudaMemcpy(o.data(), out, BH * BW * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int rr = 0; rr < BH; ++rr)
        for (int cc = 0; cc < BW; ++cc) {
          const int gx = c[0] + cc, gy = c[1] + rr;
          const float want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[(static_cast<size_t>(c[2]) * H + gy) * W + gx] : 0.f;
          if (o[rr * BW + cc] != want) ++bad;
        }
      printf("variant %s coords (%d,%d,%d): %d mismatches of %d\n", variant == 0 ? "param" : "global", c[0], c[1], c[2], bad, BH * BW);
    }
  return 0;
}
