// Weight packing + the generic row-major linear layer  Y[M,N] = X[M,K] . W[N,K]^T + b  on tcgen05.
//
// Used for (a) the low-resolution SPADE style projections (W_s . feature_maps
// is linear, so it commutes with the bilinear up-sample and runs at render resolution), and
// (b) as the smallest complete user of the tensor-core primitives in umma.cuh (self-test target).
//
// Precision modes: passes == 1  plain bf16 operands, fp32 accumulate;
//                  passes == 3  bf16x3 split (A_hi.B_hi + A_lo.B_hi + A_hi.B_lo), ~2^-16 relative,
//                               the mode that meets the 1e-3-of-fp32 parity contract.
#include "common.cuh"
#include "umma.cuh"

namespace hg {

// ------------------------------------------------------------------------------------------
// Packed weight image:  [nblocks][kchunks][part: hi, lo][Nb x 64 bf16, K-major SW128]
// Every [Nb x 64] tile is Nb*128 contiguous bytes = the exact shared-memory image a single
// cp.async.bulk drops next to the A operand.
// ------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ W, int N, int K, int ldw,
                                   const float* __restrict__ scale_ptr, float scale, int Nb, int nblocks,
                                   int kchunks, uint8_t* __restrict__ out) {
  const long total = static_cast<long>(nblocks) * Nb * kchunks * 8;
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int k8 = static_cast<int>(idx % (kchunks * 8));
  const int n = static_cast<int>(idx / (kchunks * 8));
  const int nb = n / Nb, r = n % Nb, kc = k8 / 8, c = k8 % 8;
  const float s = scale_ptr ? scale * scale_ptr[0] : scale;
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = k8 * 8 + i;
    x[i] = (n < N && k < K) ? W[static_cast<long>(n) * ldw + k] * s : 0.f;
  }
  const size_t tile_bytes = static_cast<size_t>(Nb) * 128;
  uint8_t* hi = out + (static_cast<size_t>(nb * kchunks + kc) * 2 + 0) * tile_bytes;
  uint8_t* lo = hi + tile_bytes;
  uint32_t h4[4], l4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_bf16x2(x[2 * i], x[2 * i + 1], h4[i], l4[i]);
  const uint32_t off = sw128_offset(r, c * 8);
  *reinterpret_cast<uint4*>(hi + off) = make_uint4(h4[0], h4[1], h4[2], h4[3]);
  *reinterpret_cast<uint4*>(lo + off) = make_uint4(l4[0], l4[1], l4[2], l4[3]);
}

// ------------------------------------------------------------------------------------------
// linear kernel
// ------------------------------------------------------------------------------------------
constexpr int kLinThreads = 320;  // warps 0-7: load X / epilogue, warp 8: MMA issuer, warp 9: weight producer
constexpr int kLinStages = 3;
constexpr uint32_t kChunkBytesA = 128 * 128;  // one [128 x 64] bf16 tile
constexpr uint32_t kStageBytesB = 256 * 128;  // one [256 x 64] bf16 tile (max)
constexpr uint32_t kLinSmem = 8 * kChunkBytesA + kLinStages * kStageBytesB + 256 + 1024;

template <int kPasses>
__global__ void __launch_bounds__(kLinThreads, 1)
linear_kernel(const float* __restrict__ X, int ldx, int M, int K, const uint8_t* __restrict__ Wimg, int Nb,
              int nblocks, int N, const float* __restrict__ bias, float* __restrict__ Y, int ldy) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_hi = smem;
  uint8_t* a_lo = smem + 4 * kChunkBytesA;
  uint8_t* b_st = smem + 8 * kChunkBytesA;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_st + kLinStages * kStageBytesB);
  uint64_t* a_full = bars + 0;
  uint64_t* a_empty = bars + 1;
  uint64_t* b_full = bars + 2;                  // [kLinStages]
  uint64_t* b_empty = bars + 2 + kLinStages;    // [kLinStages]
  uint64_t* acc_full = bars + 2 + 2 * kLinStages;   // [2]
  uint64_t* acc_empty = bars + 4 + 2 * kLinStages;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6 + 2 * kLinStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kchunks = (K + 63) / 64;
  const int num_tiles = (M + 127) / 128;

  if (threadIdx.x == 0) {
    mbar_init(a_full, 256);
    mbar_init(a_empty, 1);
    for (int i = 0; i < kLinStages; ++i) {
      mbar_init(b_full + i, 1);
      mbar_init(b_empty + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(acc_full + i, 1);
      mbar_init(acc_empty + i, 8);
    }
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 8) {
    // ------------------------------------------------------------ X loader + epilogue
    uint32_t acc_use = 0, it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      mbar_wait(a_empty, (it & 1) ^ 1);
      for (int r = warp; r < 128; r += 8) {
        const long grow = static_cast<long>(tile) * 128 + r;
        for (int kb = lane * 8; kb < kchunks * 64; kb += 256) {
          float x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = 0.f;
          if (grow < M) {
            if (kb + 8 <= K && (ldx & 3) == 0) {
              const float4 v0 = *reinterpret_cast<const float4*>(X + grow * ldx + kb);
              const float4 v1 = *reinterpret_cast<const float4*>(X + grow * ldx + kb + 4);
              x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w;
              x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (kb + i < K) x[i] = X[grow * ldx + kb + i];
            }
          }
          const int kc = kb >> 6;
          store_a8<kPasses == 3>(a_hi + kc * kChunkBytesA, a_lo + kc * kChunkBytesA, r, kb & 63, x);
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(a_full);

      const int q = warp & 3, h = warp >> 2;
      const int row = q * 32 + lane;
      const long grow = static_cast<long>(tile) * 128 + row;
      for (int nb = 0; nb < nblocks; ++nb, ++acc_use) {
        const uint32_t buf = acc_use & 1;
        mbar_wait(acc_full + buf, (acc_use >> 1) & 1);
        tc_fence_after();
        for (int c0 = h * 32; c0 < Nb; c0 += 64) {
          uint32_t v[32];
          tmem_ld32(tmem + (static_cast<uint32_t>(q * 32) << 16) + buf * 256 + c0, v);
          tmem_ld_wait();
          if (grow < M) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = nb * Nb + c0 + j;
              if (n < N) Y[grow * ldy + n] = __uint_as_float(v[j]) + (bias ? bias[n] : 0.f);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty + buf);
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------ MMA issuer
    {      // the warp walks the loops, one elected lane issues (umma.cuh: elect_one_sync)
      const bool leader = elect_one_sync();
      const uint32_t idesc = umma_idesc_bf16(128, Nb);
      uint32_t st = 0, ph = 0, acc_use = 0, it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        mbar_wait(a_full, it & 1);
        tc_fence_after();
        for (int nb = 0; nb < nblocks; ++nb, ++acc_use) {
          const uint32_t buf = acc_use & 1;
          mbar_wait(acc_empty + buf, ((acc_use >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d = tmem + buf * 256;
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait(b_full + st, ph);
            tc_fence_after();
            umma_k64_if(leader, d, smem_u32(a_hi + kc * kChunkBytesA), smem_u32(b_st + st * kStageBytesB), idesc, kc > 0);
            if (kPasses == 3)
              umma_k64_if(leader, d, smem_u32(a_lo + kc * kChunkBytesA), smem_u32(b_st + st * kStageBytesB), idesc, true);
            umma_commit_if(leader, b_empty + st);
            if (++st == kLinStages) { st = 0; ph ^= 1; }
            if (kPasses == 3) {
              mbar_wait(b_full + st, ph);
              tc_fence_after();
              umma_k64_if(leader, d, smem_u32(a_hi + kc * kChunkBytesA), smem_u32(b_st + st * kStageBytesB), idesc, true);
              umma_commit_if(leader, b_empty + st);
              if (++st == kLinStages) { st = 0; ph ^= 1; }
            }
          }
          umma_commit_if(leader, acc_full + buf);
        }
        umma_commit_if(leader, a_empty);
      }
    }
  } else {
    // ------------------------------------------------------------ weight producer (bulk copies from L2)
    if (lane == 0) {
      const uint32_t tile_bytes = static_cast<uint32_t>(Nb) * 128;
      uint32_t st = 0, ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int nb = 0; nb < nblocks; ++nb)
          for (int kc = 0; kc < kchunks; ++kc)
            for (int part = 0; part < (kPasses == 3 ? 2 : 1); ++part) {
              mbar_wait_backoff(b_empty + st, ph ^ 1);
              mbar_arrive_expect_tx(b_full + st, tile_bytes);
              bulk_g2s(b_st + st * kStageBytesB,
                       Wimg + (static_cast<size_t>(nb * kchunks + kc) * 2 + part) * tile_bytes, tile_bytes,
                       b_full + st);
              if (++st == kLinStages) { st = 0; ph ^= 1; }
            }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tmem);
}

}  // namespace hg

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

size_t hg_packed_weight_bytes(int N, int K, int Nb) {
  if (N <= 0 || K <= 0 || Nb <= 0) return 0;
  const size_t nblocks = (N + Nb - 1) / Nb, kchunks = (K + 63) / 64;
  return nblocks * kchunks * 2 * static_cast<size_t>(Nb) * 128;
}

int hg_pack_weight(const float* W, int N, int K, int ldw, const float* scale_dev, float scale, int Nb,
                   void* out_img, size_t out_bytes, void* stream) {
  HG_REQUIRE(W && out_img, "hg_pack_weight: null pointer");
  HG_REQUIRE(Nb >= 16 && Nb <= 256 && Nb % 16 == 0, "hg_pack_weight: Nb=%d must be a multiple of 16 in [16,256]", Nb);
  HG_REQUIRE(N > 0 && K > 0 && ldw >= K, "hg_pack_weight: bad shape N=%d K=%d ldw=%d", N, K, ldw);
  HG_REQUIRE(out_bytes >= hg_packed_weight_bytes(N, K, Nb), "hg_pack_weight: output buffer too small");
  HG_REQUIRE((reinterpret_cast<uintptr_t>(out_img) & 15) == 0, "hg_pack_weight: output must be 16-byte aligned");
  const int nblocks = (N + Nb - 1) / Nb, kchunks = (K + 63) / 64;
  const long total = static_cast<long>(nblocks) * Nb * kchunks * 8;
  hg::pack_weight_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      W, N, K, ldw, scale_dev, scale, Nb, nblocks, kchunks, static_cast<uint8_t*>(out_img));
  return hg::check_launch("hg_pack_weight");
}

int hg_linear(const float* X, int ldx, int M, int K, const void* Wimg, int Nb, int N, const float* bias, float* Y,
              int ldy, int passes, void* stream) {
  HG_REQUIRE(X && Wimg && Y, "hg_linear: null pointer");
  HG_REQUIRE(M > 0 && K > 0 && K <= 256 && N > 0, "hg_linear: need 0 < K <= 256 (got %d), M=%d N=%d", K, M, N);
  HG_REQUIRE(Nb >= 16 && Nb <= 256 && Nb % 16 == 0, "hg_linear: Nb=%d must be a multiple of 16 in [16,256]", Nb);
  HG_REQUIRE(passes == 1 || passes == 3, "hg_linear: passes must be 1 (bf16) or 3 (bf16x3)");
  HG_REQUIRE(ldx >= K && ldy >= N, "hg_linear: leading dimensions too small");
  const int nblocks = (N + Nb - 1) / Nb;
  const int num_tiles = (M + 127) / 128;
  const int grid = num_tiles < hg::num_sms() ? num_tiles : hg::num_sms();
  auto st = static_cast<cudaStream_t>(stream);
  const auto* img = static_cast<const uint8_t*>(Wimg);
  cudaError_t e;
  if (passes == 3) {
    e = cudaFuncSetAttribute(hg::linear_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, hg::kLinSmem);
    if (e != cudaSuccess) { hg::set_error("hg_linear: smem opt-in failed: %s", cudaGetErrorString(e)); return 2; }
    hg::linear_kernel<3><<<grid, hg::kLinThreads, hg::kLinSmem, st>>>(X, ldx, M, K, img, Nb, nblocks, N, bias, Y, ldy);
  } else {
    e = cudaFuncSetAttribute(hg::linear_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, hg::kLinSmem);
    if (e != cudaSuccess) { hg::set_error("hg_linear: smem opt-in failed: %s", cudaGetErrorString(e)); return 2; }
    hg::linear_kernel<1><<<grid, hg::kLinThreads, hg::kLinSmem, st>>>(X, ldx, M, K, img, Nb, nblocks, N, bias, Y, ldy);
  }
  return hg::check_launch("hg_linear");
}

}  // extern "C"
