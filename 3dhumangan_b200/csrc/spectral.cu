// Spectral normalisation of a whole network's convolution weights in ONE launch.
//
// torch.nn.utils.spectral_norm (applied at lib/components/map3d_layers.py:205-206 to the 18 synthesis convolutions and
// at lib/discriminators/unet_discriminators.py:18 to the 30 discriminator convolutions) runs, in front of every TRAINING
// forward and per layer:   v <- normalize(W^T u),  u <- normalize(W v),  sigma = u . (W v),  weight = W / sigma
// with W = weight_orig viewed as [N, K], normalize(x) = x / max(|x|_2, eps), eps = 1e-12, buffers u / v updated in place;
// in eval mode only sigma = u . (W v) with the stored vectors.  That is 5-8 tiny launches per layer (48 layers, 3-4 forwards
// per iteration) in the reference.  Here: one CTA per matrix, a table of matrices per launch, 1/sigma written to a vector
// that the operand-packing kernel (hg_pack_weight) and the training graph consume.
//
// Memory bound and tiny (the largest matrix, 512 x 4608 fp32, is read twice); deterministic: every sum is evaluated in a
// fixed order (no atomics), so repeated runs and CUDA-graph replays give identical u, v, sigma.
#include "common.cuh"

namespace hg {

struct SnEntry {
  const float* w;   // [N, K] row-major
  float* u;         // [N]
  float* v;         // [K]
  int N, K;
};

constexpr int kSnThreads = 1024;

__device__ __forceinline__ float block_sum(float x, float* red) {      // fixed-order tree: deterministic
  for (int s = 16; s > 0; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = x;
  __syncthreads();
  float t = (threadIdx.x < (kSnThreads >> 5)) ? red[threadIdx.x] : 0.f;
  if (warp == 0) {
    for (int s = 16; s > 0; s >>= 1) t += __shfl_xor_sync(0xffffffffu, t, s);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// dynamic shared memory: us [N] | vs [K] | part [groups x Kp] (only when K < 1024)
__global__ void __launch_bounds__(kSnThreads) spectral_kernel(const SnEntry* __restrict__ table, float* __restrict__ inv_sigma,
                                                              int training, float eps) {
  extern __shared__ float sn_smem[];
  __shared__ float red[33];
  const SnEntry e = table[blockIdx.x];
  const int N = e.N, K = e.K;
  float* us = sn_smem;
  float* vs = us + N;
  float* part = vs + K;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < N; i += kSnThreads) us[i] = e.u[i];
  for (int i = tid; i < K; i += kSnThreads) vs[i] = e.v[i];
  __syncthreads();

  if (training) {
    // ---- t = W^T u: a thread owns columns (coalesced across the warp); row slices per thread group when K is small
    const int Kp = (K + 31) & ~31;
    const int groups = Kp >= kSnThreads ? 1 : kSnThreads / Kp;
    if (groups == 1) {
      for (int k = tid; k < K; k += kSnThreads) {
        float acc = 0.f;
        const float* col = e.w + k;
#pragma unroll 8
        for (int n = 0; n < N; ++n) acc = fmaf(__ldg(col + static_cast<long>(n) * K), us[n], acc);
        vs[k] = acc;
      }
    } else {
      const int g = tid / Kp, k = tid - g * Kp;
      if (g < groups && k < K) {
        const int per = (N + groups - 1) / groups;
        const int n0 = g * per, n1 = min(N, n0 + per);
        float acc = 0.f;
        const float* col = e.w + k;
#pragma unroll 8
        for (int n = n0; n < n1; ++n) acc = fmaf(__ldg(col + static_cast<long>(n) * K), us[n], acc);
        part[g * Kp + k] = acc;
      }
      __syncthreads();
      for (int kk = tid; kk < K; kk += kSnThreads) {
        float acc = 0.f;
        for (int gg = 0; gg < groups; ++gg) acc += part[gg * Kp + kk];
        vs[kk] = acc;
      }
    }
    __syncthreads();
    float sq = 0.f;
    for (int k = tid; k < K; k += kSnThreads) sq = fmaf(vs[k], vs[k], sq);
    const float nv = sqrtf(block_sum(sq, red));
    const float rv = 1.f / fmaxf(nv, eps);
    for (int k = tid; k < K; k += kSnThreads) {
      const float val = vs[k] * rv;
      vs[k] = val;
      e.v[k] = val;
    }
    __syncthreads();
  }

  // ---- s = W v: a warp owns rows, lanes stride the columns (coalesced), shuffle tree
  for (int n = warp; n < N; n += (kSnThreads >> 5)) {
    const float* row = e.w + static_cast<long>(n) * K;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(__ldg(row + k), vs[k], acc);
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) part[n] = acc;          // part is at least N floats long (host sizes it)
  }
  __syncthreads();
  float sigma;
  if (training) {
    float sq = 0.f;
    for (int n = tid; n < N; n += kSnThreads) sq = fmaf(part[n], part[n], sq);
    const float ss = block_sum(sq, red);
    const float ru = 1.f / fmaxf(sqrtf(ss), eps);
    for (int n = tid; n < N; n += kSnThreads) e.u[n] = part[n] * ru;
    sigma = ss * ru;                       // u . (W v) with u = (W v) / max(|W v|, eps)
  } else {
    float d = 0.f;
    for (int n = tid; n < N; n += kSnThreads) d = fmaf(us[n], part[n], d);
    sigma = block_sum(d, red);
  }
  if (tid == 0) inv_sigma[blockIdx.x] = 1.f / sigma;
}

}  // namespace hg

extern "C" {

// Layout of one table entry as the host writes it (5 x 8 bytes): w, u, v pointers, then N and K as int32 + padding.
int hg_spectral_entry_bytes(void) { return static_cast<int>(sizeof(hg::SnEntry)); }

int hg_spectral_norm(const void* table, int count, int max_n, int max_k, float* inv_sigma, int training, float eps,
                     void* stream) {
  HG_REQUIRE(table && inv_sigma && count > 0, "hg_spectral_norm: bad arguments");
  HG_REQUIRE(max_n > 0 && max_k > 0, "hg_spectral_norm: bad matrix bounds");
  const int kp = (max_k + 31) & ~31;
  const int groups = kp >= hg::kSnThreads ? 1 : hg::kSnThreads / kp;
  (void)groups;
  size_t part = hg::kSnThreads;             // any entry with K < 1024 uses groups x Kp <= 1024 partial sums
  if (part < static_cast<size_t>(max_n)) part = max_n;
  const size_t smem = (static_cast<size_t>(max_n) + max_k + part) * sizeof(float);
  HG_REQUIRE(smem <= 200 * 1024, "hg_spectral_norm: matrix too large for one CTA (N=%d K=%d)", max_n, max_k);
  cudaError_t e = cudaFuncSetAttribute(hg::spectral_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) { hg::set_error("hg_spectral_norm: smem opt-in failed: %s", cudaGetErrorString(e)); return 2; }
  hg::spectral_kernel<<<count, hg::kSnThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const hg::SnEntry*>(table), inv_sigma, training, eps);
  return hg::check_launch("hg_spectral_norm");
}

}  // extern "C"
