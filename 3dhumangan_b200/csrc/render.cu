// Fused pose-mapping renderer: per-point FiLM-SIREN MLP + alpha compositing along each ray.
//
// Replaces (reference file:line):
//   COORDCONCATSIREN.forward        lib/implicit_funcitions/modulated.py:41-75
//   SineLayer / FiLMLayer           lib/components/pigan_layers.py:63-87
//   vr.ray_integration              lib/generators/volume_rendering.py:12-56
//   the [B,N,F+4] / [B,N,256..512] intermediates of Map3DGenerator.render (map3d_generator.py:427-515)
//
// A CTA owns tiles of 128 sample points = (128/S) whole rays.  The eight GEMM steps of the MLP run
// back to back on tcgen05 with the activations never leaving the SM: accumulator (TMEM) ->
// sin(F*acc + P) in registers -> bf16 hi/lo operand in shared memory -> next tcgen05.mma, while the
// weight tiles stream from L2 through the bulk-copy engine.  FiLM frequency/phase, bias and the
// x30 of the sine layers are folded into one (F, P) table per layer and sample (host side).  The
// sigma and rgb heads are dot products on the fp32 activations; transmittance is a per-ray scan in
// shared memory and the weighted feature sum is a warp shuffle transpose-reduce, so that only
// [rays, 256 feat + 3 rgb + depth] ever reaches HBM.
//
// Layer schedule per tile (acc A = TMEM cols 0..255, acc B = cols 256..511):
//   L0   rec[K=64: xyz*s, geo31, 0]  x W01(coord | geo)      -> A (coord pre-act), B (geo pre-act)
//   E0   a = sin(30*(A+b))                                    -> operand
//   L1a  a x Wn0[:, 0:256]                                    -> A
//   E1   g = sin(30*(B+b))                                    -> operand
//   L1b  g x Wn0[:, 256:512]                                  -> A (accumulate)
//   E2..E4  x = sin(F*(acc+b)+P) ; L: network.1..3            -> B, A, B
//   E5   x4 (+ sigma head) ; L: color_layer_sine[:, 3:]       -> A       (view-direction term in P)
//   E6   c (+ rgb head)    ; L: feature_layer_linear          -> B
//   E7   feat + compositing
#include "common.cuh"
#include "umma.cuh"

namespace hg {

constexpr int kRH = 256;           // hidden_dim == feature_dim
constexpr int kRenThreads = 320;   // warps 0-7 rows, 8 MMA, 9 weight producer
constexpr int kRenStages = 2;
constexpr uint32_t kRA = 128 * 128;   // operand chunk [128 x 64] bf16
constexpr uint32_t kRB = 256 * 128;   // weight stage  [256 x 64] bf16
constexpr int kFilmLayers = 7;        // coord, geo, network.0..3, color
constexpr int kWeightStages = 60;     // per tile, hi+lo (see schedule above)
constexpr int kRayOut = 260;          // floats per ray: 256 feat, 3 rgb, depth

struct RenderArgs {
  const float* rec;      // [B,N,36] point records (geo.cu)
  const float* z_vals;   // [B,N]
  const float* noise;    // [B,N] N(0,1) draws or null
  const float* film;     // [B,7,2,256]: per layer F (multiplier) and P (offset): x = sin(F*acc + P)
  const uint8_t* wblob;  // 60 packed weight stages in schedule order
  const float* w_sigma;  // [256]
  const float* w_rgb;    // [3,256]
  const float* b_feat;   // [256]
  const float* heads_b;  // [4]: b_sigma, b_rgb[3]
  float* ray_out;        // [B,R,260]
  float* weights_out;    // [B,N] or null
  float* raw_out;        // [B,N,260] per-point (rgb 3, feat 256, sigma) instead of compositing, or null
  int B, R, S;           // rays per image, samples per ray (power of two <= 128)
  float noise_std;
  int white_back, last_back, clamp_softplus;
};

struct RenSmem {
  uint8_t* a_hi;
  uint8_t* a_lo;
  uint8_t* b_st;
  float* film;     // [7][2][256]
  float* w_sigma;  // [256]
  float* w_rgb;    // [3][256]
  float* b_feat;   // [256]
  float* part;     // [2][128][4] sigma / rgb partial dots per column half
  float* tr;       // [128] 1 - alpha + 1e-12
  float* wgt;      // [128] compositing weights
  float* zs;       // [128]
  float* rayw;     // [128] per-ray sum of weights (first rays_per_tile entries)
  uint64_t* bars;
  uint32_t* tmem_slot;
};
constexpr uint32_t kRenFloats = kFilmLayers * 2 * kRH + kRH + 3 * kRH + kRH + 2 * 128 * 4 + 4 * 128;
constexpr uint32_t kRenSmemBytes = 8 * kRA + kRenStages * kRB + kRenFloats * 4 + 16 * 8 + 16 + 1024;

__device__ __forceinline__ RenSmem ren_carve(uint8_t* raw) {
  uint8_t* s = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  RenSmem m;
  m.a_hi = s;
  m.a_lo = s + 4 * kRA;
  m.b_st = s + 8 * kRA;
  float* f = reinterpret_cast<float*>(m.b_st + kRenStages * kRB);
  m.film = f; f += kFilmLayers * 2 * kRH;
  m.w_sigma = f; f += kRH;
  m.w_rgb = f; f += 3 * kRH;
  m.b_feat = f; f += kRH;
  m.part = f; f += 2 * 128 * 4;
  m.tr = f; f += 128;
  m.wgt = f; f += 128;
  m.zs = f; f += 128;
  m.rayw = f; f += 128;
  m.bars = reinterpret_cast<uint64_t*>(f);
  m.tmem_slot = reinterpret_cast<uint32_t*>(m.bars + 16);
  return m;
}
enum { RA_FULL = 0 /*4*/, RB_FULL = 4 /*2*/, RB_EMPTY = 6 /*2*/, RL_FULL = 8 };

__device__ __forceinline__ void ren_rows_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// sin(t) for |t| up to a few thousand: two-term Cody-Waite reduction by 2*pi, then the SFU sine on
// [-pi, pi] (abs error ~2^-21).  The reference path evaluates torch.sin on fp32 tensors.
__device__ __forceinline__ float sin_reduced(float t) {
  const float y = t * 0.15915494309189535f;
  const float k = (y + 12582912.f) - 12582912.f;  // round to nearest integer (|y| < 2^22)
  float r = fmaf(-k, 6.2831854820251465f, t);
  r = fmaf(-k, -1.7484555314695172e-07f, r);
  return __sinf(r);
}

// the same evaluation on a pair (packed fp32 for everything in front of the SFU: identical operations, half the instructions)
__device__ __forceinline__ float2 sin_reduced2(float2 t) {
  const float2 y = __fmul2_rn(t, make_float2(0.15915494309189535f, 0.15915494309189535f));
  const float2 big = make_float2(12582912.f, 12582912.f), nbig = make_float2(-12582912.f, -12582912.f);
  const float2 k = __fadd2_rn(__fadd2_rn(y, big), nbig);
  float2 r = __ffma2_rn(k, make_float2(-6.2831854820251465f, -6.2831854820251465f), t);
  r = __ffma2_rn(k, make_float2(1.7484555314695172e-07f, 1.7484555314695172e-07f), r);
  return make_float2(__sinf(r.x), __sinf(r.y));
}

__device__ __forceinline__ float transpose_reduce32r(float (&v)[32], int lane) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const bool upper = (lane & w) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float send = upper ? v[i] : v[i + w];
      const float keep = upper ? v[i + w] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  return v[0];
}

template <int kPasses>
__device__ __forceinline__ void ren_mma_chunk(const RenSmem& m, bool leader, uint32_t& st, uint32_t& ph, uint32_t tmem_d,
                                              uint32_t a_hi, uint32_t a_lo, uint32_t idesc, bool accumulate) {
  mbar_wait(m.bars + RB_FULL + st, ph);
  tc_fence_after();
  umma_k64_if(leader, tmem_d, a_hi, smem_u32(m.b_st + st * kRB), idesc, accumulate);
  if (kPasses == 3) umma_k64_if(leader, tmem_d, a_lo, smem_u32(m.b_st + st * kRB), idesc, true);
  umma_commit_if(leader, m.bars + RB_EMPTY + st);
  if (++st == kRenStages) { st = 0; ph ^= 1; }
  if (kPasses == 3) {
    mbar_wait(m.bars + RB_FULL + st, ph);
    tc_fence_after();
    umma_k64_if(leader, tmem_d, a_hi, smem_u32(m.b_st + st * kRB), idesc, true);
    umma_commit_if(leader, m.bars + RB_EMPTY + st);
    if (++st == kRenStages) { st = 0; ph ^= 1; }
  }
}

// Activation epilogue of one layer: TMEM accumulator -> x = sin(F*acc + P) -> bf16 hi/lo operand.
// kHead: 0 none, 1 sigma dot (1 value), 2 rgb dots (3 values) accumulated into part[h][row][*].
template <int kPasses, int kHead>
__device__ __forceinline__ void film_epilogue(const RenSmem& m, uint32_t tmem_acc, int layer, int warp, int lane,
                                              bool defer_arrive) {
  const int q = warp & 3, h = warp >> 2;
  const int row = q * 32 + lane;
  uint32_t F = smem_u32(m.film + (layer * 2 + 0) * kRH);
  uint32_t P = smem_u32(m.film + (layer * 2 + 1) * kRH);
  opaque(F);   // the per-sample FiLM table is refreshed between tiles: table loads stay inside this call
  opaque(P);
  uint32_t wsig = smem_u32(m.w_sigma), wrgb = smem_u32(m.w_rgb);
  opaque(wsig);
  opaque(wrgb);
  float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll 1
  for (int kc = 0; kc < 4; ++kc) {
    const int c0 = kc * 64 + h * 32;
    uint32_t raw[32];
    tmem_ld32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + c0, raw);
    tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float x[8], f8[8], p8[8];
      lds8(F + (c0 + g * 8) * 4, f8);
      lds8(P + (c0 + g * 8) * 4, p8);
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const float2 s2 = sin_reduced2(__ffma2_rn(make_float2(f8[j], f8[j + 1]),
                                                  make_float2(__uint_as_float(raw[g * 8 + j]), __uint_as_float(raw[g * 8 + j + 1])),
                                                  make_float2(p8[j], p8[j + 1])));
        x[j] = s2.x;
        x[j + 1] = s2.y;
      }
      if (kHead == 1) {
        float w8[8];
        lds8(wsig + (c0 + g * 8) * 4, w8);
#pragma unroll
        for (int j = 0; j < 8; ++j) d0 = fmaf(x[j], w8[j], d0);
      }
      if (kHead == 2) {
        float w0[8], w1[8], w2[8];
        lds8(wrgb + (c0 + g * 8) * 4, w0);
        lds8(wrgb + (kRH + c0 + g * 8) * 4, w1);
        lds8(wrgb + (2 * kRH + c0 + g * 8) * 4, w2);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          d0 = fmaf(x[j], w0[j], d0);
          d1 = fmaf(x[j], w1[j], d1);
          d2 = fmaf(x[j], w2[j], d2);
        }
      }
      store_a8<kPasses == 3>(m.a_hi + kc * kRA, m.a_lo + kc * kRA, row, h * 32 + g * 8, x);
    }
    if (!defer_arrive) {
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(m.bars + RA_FULL + kc);
    }
  }
  if (defer_arrive) {
    tc_fence_before();
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0)
      for (int kc = 0; kc < 4; ++kc) mbar_arrive(m.bars + RA_FULL + kc);
  }
  if (kHead == 1) m.part[(h * 128 + row) * 4 + 0] = d0;
  if (kHead == 2) {
    m.part[(h * 128 + row) * 4 + 1] = d0;
    m.part[(h * 128 + row) * 4 + 2] = d1;
    m.part[(h * 128 + row) * 4 + 3] = d2;
  }
}

template <int kPasses>
__global__ void __launch_bounds__(kRenThreads, 1) render_mlp_kernel(RenderArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const RenSmem m = ren_carve(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kRH; i += blockDim.x) {
    m.w_sigma[i] = a.w_sigma[i];
    m.b_feat[i] = a.b_feat[i];
  }
  for (int i = threadIdx.x; i < 3 * kRH; i += blockDim.x) m.w_rgb[i] = a.w_rgb[i];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(m.bars + RA_FULL + i, 8);
    for (int i = 0; i < kRenStages; ++i) {
      mbar_init(m.bars + RB_FULL + i, 1);
      mbar_init(m.bars + RB_EMPTY + i, 1);
    }
    mbar_init(m.bars + RL_FULL, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<512>(m.tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *m.tmem_slot;
  const uint32_t accA = tmem, accB = tmem + 256;

  const int S = a.S;
  const int rpt = 128 / S;                               // rays per tile
  const int tiles_per_img = (a.R + rpt - 1) / rpt;
  const int num_tiles = a.B * tiles_per_img;
  const int my_tiles = (num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const long N = static_cast<long>(a.R) * S;

  if (warp < 8) {
    const int q = warp & 3, h = warp >> 2;
    const int row = q * 32 + lane;
    const int rl = row / S, s = row % S;                 // ray within the tile, sample along the ray
    uint32_t lph = 0;                                    // RL_FULL phase
    int cur_b = -1;
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int b = tile / tiles_per_img;
      const int ray0 = (tile % tiles_per_img) * rpt;
      const int ray = ray0 + rl;
      const bool valid = ray < a.R;
      const long gp = static_cast<long>(b) * N + static_cast<long>(ray) * S + s;
      if (b != cur_b) {
        ren_rows_barrier();
        for (int i = threadIdx.x; i < kFilmLayers * 2 * kRH; i += 256)
          m.film[i] = a.film[static_cast<long>(b) * kFilmLayers * 2 * kRH + i];
        cur_b = b;
        ren_rows_barrier();
      }
      // ---- A0: point record -> operand chunk 0 (K = 64: 36 values + zeros)
      {
        float x[8];
        const float4* rp = reinterpret_cast<const float4*>(a.rec + gp * 36);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int k0 = h * 32 + g * 8;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int f4 = (k0 >> 2) + u;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid && f4 < 9) v = __ldg(rp + f4);
            x[u * 4 + 0] = v.x; x[u * 4 + 1] = v.y; x[u * 4 + 2] = v.z; x[u * 4 + 3] = v.w;
          }
          store_a8<kPasses == 3>(m.a_hi, m.a_lo, row, k0, x);
        }
        if (h == 0) m.zs[row] = (valid && a.z_vals) ? a.z_vals[gp] : 0.f;
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(m.bars + RA_FULL + 0);
      }
      // ---- E0 .. E6
      mbar_wait_sleep(m.bars + RL_FULL, lph); lph ^= 1; tc_fence_after();
      film_epilogue<kPasses, 0>(m, accA, 0, warp, lane, /*defer=*/true);
      mbar_wait_sleep(m.bars + RL_FULL, lph); lph ^= 1; tc_fence_after();
      film_epilogue<kPasses, 0>(m, accB, 1, warp, lane, false);
      mbar_wait_sleep(m.bars + RL_FULL, lph); lph ^= 1; tc_fence_after();
      film_epilogue<kPasses, 0>(m, accA, 2, warp, lane, false);
      mbar_wait_sleep(m.bars + RL_FULL, lph); lph ^= 1; tc_fence_after();
      film_epilogue<kPasses, 0>(m, accB, 3, warp, lane, false);
      mbar_wait_sleep(m.bars + RL_FULL, lph); lph ^= 1; tc_fence_after();
      film_epilogue<kPasses, 0>(m, accA, 4, warp, lane, false);
      mbar_wait_sleep(m.bars + RL_FULL, lph); lph ^= 1; tc_fence_after();
      film_epilogue<kPasses, 1>(m, accB, 5, warp, lane, false);
      mbar_wait_sleep(m.bars + RL_FULL, lph); lph ^= 1; tc_fence_after();
      film_epilogue<kPasses, 2>(m, accA, 6, warp, lane, false);

      if (a.raw_out) {
        // ---- per-point outputs of COORDCONCATSIREN.forward (modulated.py:70-73): [rgb, feat, sigma]
        ren_rows_barrier();
        float* po = a.raw_out + gp * kRayOut;
        if (h == 0 && valid) {
          po[259] = m.part[row * 4] + m.part[(128 + row) * 4] + a.heads_b[0];
          for (int j = 0; j < 3; ++j) {
            const float dot = m.part[row * 4 + 1 + j] + m.part[(128 + row) * 4 + 1 + j] + a.heads_b[1 + j];
            po[j] = 1.f / (1.f + expf(-dot));
          }
        }
        mbar_wait_sleep(m.bars + RL_FULL, lph); lph ^= 1; tc_fence_after();
#pragma unroll 1
        for (int kc = 0; kc < 4; ++kc) {
          const int c0 = kc * 64 + h * 32;
          uint32_t raw[32];
          tmem_ld32(accB + (static_cast<uint32_t>(q * 32) << 16) + c0, raw);
          tmem_ld_wait();
          if (valid)
            for (int j = 0; j < 32; ++j) po[3 + c0 + j] = __uint_as_float(raw[j]) + m.b_feat[c0 + j];
        }
        tc_fence_before();
        ren_rows_barrier();
        continue;
      }
      // ---- compositing weights (volume_rendering.py:12-38); overlaps the feature GEMM
      ren_rows_barrier();
      float alpha = 0.f;
      if (h == 0) {
        const float sigma = m.part[row * 4] + m.part[(128 + row) * 4] + a.heads_b[0];
        const float delta = (s == S - 1) ? 1e9f : m.zs[row + 1] - m.zs[row];
        float pre = sigma;
        if (a.noise) pre += (valid ? a.noise[gp] : 0.f) * a.noise_std;
        const float dens = a.clamp_softplus ? (pre > 20.f ? pre : log1pf(expf(pre))) : fmaxf(pre, 0.f);
        alpha = 1.f - expf(-delta * dens);
        m.tr[row] = 1.f - alpha + 1e-12f;
      }
      ren_rows_barrier();
      if (h == 0) {
        float T = 1.f;
        for (int k = 0; k < s; ++k) T *= m.tr[rl * S + k];
        m.wgt[row] = valid ? alpha * T : 0.f;
      }
      ren_rows_barrier();
      if (h == 0 && s == 0) {
        float W = 0.f;
        for (int k = 0; k < S; ++k) W += m.wgt[rl * S + k];
        m.rayw[rl] = W;
      }
      ren_rows_barrier();
      const float Wsum = m.rayw[rl];
      float w = m.wgt[row];
      const float w_depth = w + ((s == S - 1) ? 1.f - Wsum : 0.f);
      if (a.last_back) w = w_depth;
      if (a.weights_out && h == 0 && valid) a.weights_out[gp] = w;
      const float back = a.white_back ? 1.f - Wsum : 0.f;

      // ---- E7: feature accumulator -> weighted sum over the ray
      mbar_wait_sleep(m.bars + RL_FULL, lph); lph ^= 1; tc_fence_after();
      float* ro = a.ray_out + (static_cast<long>(b) * a.R + ray) * kRayOut;
      if (S == 32) {
        // one warp == one ray: shuffle transpose-reduce, lane j ends with column j's sum
#pragma unroll 1
        for (int kc = 0; kc < 4; ++kc) {
          const int c0 = kc * 64 + h * 32;
          uint32_t raw[32];
          tmem_ld32(accB + (static_cast<uint32_t>(q * 32) << 16) + c0, raw);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float b8[8];
            uint32_t bfa = smem_u32(m.b_feat);
            opaque(bfa);
            lds8(bfa + (c0 + g * 8) * 4, b8);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[g * 8 + j] = w * (__uint_as_float(raw[g * 8 + j]) + b8[j]);
          }
          const float tot = transpose_reduce32r(v, lane);
          if (valid) ro[c0 + lane] = tot + back;   // `ro`/`valid` are warp-uniform here (S == 32)
        }
        if (h == 0) {
          float e[4];
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float dot = m.part[row * 4 + 1 + j] + m.part[(128 + row) * 4 + 1 + j] + a.heads_b[1 + j];
            e[j] = w / (1.f + expf(-dot));
          }
          e[3] = w_depth * m.zs[row];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            for (int o = 16; o > 0; o >>= 1) e[j] += __shfl_xor_sync(0xffffffffu, e[j], o);
          if (valid && lane < 4) ro[256 + lane] = (lane == 0 ? e[0] : lane == 1 ? e[1] : lane == 2 ? e[2] : e[3]) + (lane < 3 ? back : 0.f);
        }
        tc_fence_before();
      } else {
        // generic S: stage w*feat through shared memory (aliases the operand buffer, free by now)
        float* scratch = reinterpret_cast<float*>(m.a_hi);          // [2][128][33]
        float* mine = scratch + (h * 128 + row) * 33;
#pragma unroll 1
        for (int kc = 0; kc < 4; ++kc) {
          const int c0 = kc * 64 + h * 32;
          uint32_t raw[32];
          tmem_ld32(accB + (static_cast<uint32_t>(q * 32) << 16) + c0, raw);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) mine[j] = w * (__uint_as_float(raw[j]) + m.b_feat[c0 + j]);
          ren_rows_barrier();
          // 256 threads: (ray r2, column j2) pairs; thread handles rays r2, r2 + 8, ...
          for (int r2 = warp & 3; r2 < rpt; r2 += 4) {
            float acc = 0.f;
            const float* src = scratch + (h * 128 + r2 * S) * 33 + lane;
            for (int k = 0; k < S; ++k) acc += src[k * 33];
            if (ray0 + r2 < a.R)
              a.ray_out[(static_cast<long>(b) * a.R + ray0 + r2) * kRayOut + c0 + lane] =
                  acc + (a.white_back ? 1.f - m.rayw[r2] : 0.f);
          }
          ren_rows_barrier();
        }
        if (h == 0) {
          float e[4];
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float dot = m.part[row * 4 + 1 + j] + m.part[(128 + row) * 4 + 1 + j] + a.heads_b[1 + j];
            e[j] = w / (1.f + expf(-dot));
          }
          e[3] = w_depth * m.zs[row];
#pragma unroll
          for (int j = 0; j < 4; ++j) mine[j] = e[j];
        }
        ren_rows_barrier();
        if (h == 0 && s == 0 && valid) {
          float e[4] = {0.f, 0.f, 0.f, 0.f};
          for (int k = 0; k < S; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) e[j] += scratch[(rl * S + k) * 33 + j];
          ro[256] = e[0] + back; ro[257] = e[1] + back; ro[258] = e[2] + back; ro[259] = e[3];
        }
        tc_fence_before();
        ren_rows_barrier();   // scratch is the next tile's operand buffer
      }
    }
  } else if (warp == 8) {
    {      // the warp walks the loops, one elected lane issues (umma.cuh: elect_one_sync)
      const bool leader = elect_one_sync();
      const uint32_t idesc = umma_idesc_bf16(128, 256);
      uint32_t st = 0, ph = 0;
      uint32_t aph[4] = {0, 0, 0, 0};
      auto wait_a = [&](int kc) {
        mbar_wait(m.bars + RA_FULL + kc, aph[kc]);
        aph[kc] ^= 1;
        tc_fence_after();
      };
      auto A_hi = [&](int kc) { return smem_u32(m.a_hi + kc * kRA); };
      auto A_lo = [&](int kc) { return smem_u32(m.a_lo + kc * kRA); };
      for (int it = 0; it < my_tiles; ++it) {
        // L0: coord -> A, geo -> B (both read operand chunk 0)
        wait_a(0);
        ren_mma_chunk<kPasses>(m, leader, st, ph, accA, A_hi(0), A_lo(0), idesc, false);
        ren_mma_chunk<kPasses>(m, leader, st, ph, accB, A_hi(0), A_lo(0), idesc, false);
        umma_commit_if(leader, m.bars + RL_FULL);
        // L1a: a x Wn0[:, :256] -> A
        for (int kc = 0; kc < 4; ++kc) {
          wait_a(kc);
          ren_mma_chunk<kPasses>(m, leader, st, ph, accA, A_hi(kc), A_lo(kc), idesc, kc > 0);
        }
        umma_commit_if(leader, m.bars + RL_FULL);
        // L1b: g x Wn0[:, 256:] -> A (accumulate)
        for (int kc = 0; kc < 4; ++kc) {
          wait_a(kc);
          ren_mma_chunk<kPasses>(m, leader, st, ph, accA, A_hi(kc), A_lo(kc), idesc, true);
        }
        umma_commit_if(leader, m.bars + RL_FULL);
        // network.1, .2, .3, color, feature: B, A, B, A, B
        for (int l = 0; l < 5; ++l) {
          const uint32_t acc = (l & 1) ? accA : accB;
          for (int kc = 0; kc < 4; ++kc) {
            wait_a(kc);
            ren_mma_chunk<kPasses>(m, leader, st, ph, acc, A_hi(kc), A_lo(kc), idesc, kc > 0);
          }
          umma_commit_if(leader, m.bars + RL_FULL);
        }
      }
    }
  } else {
    if (lane == 0) {
      uint32_t st = 0, ph = 0;
      for (int it = 0; it < my_tiles; ++it)
        for (int sidx = 0; sidx < kWeightStages; ++sidx) {
          if (kPasses == 1 && (sidx & 1)) continue;
          mbar_wait_backoff(m.bars + RB_EMPTY + st, ph ^ 1);
          mbar_arrive_expect_tx(m.bars + RB_FULL + st, kRB);
          bulk_g2s(m.b_st + st * kRB, a.wblob + static_cast<size_t>(sidx) * kRB, kRB, m.bars + RB_FULL + st);
          if (++st == kRenStages) { st = 0; ph ^= 1; }
        }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tmem);
}

}  // namespace hg

extern "C" {

size_t hg_render_weight_blob_bytes(void) { return static_cast<size_t>(hg::kWeightStages) * hg::kRB; }

int hg_render_mlp(const float* rec, const float* z_vals, const float* noise, const float* film, const void* wblob,
                  const float* w_sigma, const float* w_rgb, const float* b_feat, const float* heads_b, float* ray_out,
                  float* weights_out, float* raw_out, int B, int R, int S, int hidden, float noise_std, int white_back, int last_back,
                  int clamp_softplus, int passes, void* stream) {
  HG_REQUIRE(hidden == hg::kRH, "hg_render_mlp: only hidden_dim == %d is supported (got %d)", hg::kRH, hidden);
  HG_REQUIRE(rec && film && wblob && w_sigma && w_rgb && b_feat && heads_b, "hg_render_mlp: null pointer");
  HG_REQUIRE(raw_out || (ray_out && z_vals), "hg_render_mlp: need ray_out + z_vals (compositing) or raw_out (per-point)");
  HG_REQUIRE(B > 0 && R > 0, "hg_render_mlp: bad shape");
  HG_REQUIRE(S >= 2 && S <= 128 && (S & (S - 1)) == 0, "hg_render_mlp: samples per ray must be a power of two in [2,128] (got %d)", S);
  HG_REQUIRE(passes == 1 || passes == 3, "hg_render_mlp: passes must be 1 or 3");
  HG_REQUIRE((reinterpret_cast<uintptr_t>(rec) & 15) == 0 && (reinterpret_cast<uintptr_t>(wblob) & 15) == 0,
             "hg_render_mlp: rec / wblob must be 16-byte aligned");
  hg::RenderArgs a{rec, z_vals, noise, film, static_cast<const uint8_t*>(wblob), w_sigma, w_rgb, b_feat, heads_b,
                   ray_out, weights_out, raw_out, B, R, S, noise_std, white_back, last_back, clamp_softplus};
  const int rpt = 128 / S;
  const int tiles = B * ((R + rpt - 1) / rpt);
  const int grid = tiles < hg::num_sms() ? tiles : hg::num_sms();
  auto st = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (passes == 3) {
    e = cudaFuncSetAttribute(hg::render_mlp_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, hg::kRenSmemBytes);
    if (e != cudaSuccess) { hg::set_error("hg_render_mlp: smem opt-in failed: %s", cudaGetErrorString(e)); return 2; }
    hg::render_mlp_kernel<3><<<grid, hg::kRenThreads, hg::kRenSmemBytes, st>>>(a);
  } else {
    e = cudaFuncSetAttribute(hg::render_mlp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, hg::kRenSmemBytes);
    if (e != cudaSuccess) { hg::set_error("hg_render_mlp: smem opt-in failed: %s", cudaGetErrorString(e)); return 2; }
    hg::render_mlp_kernel<1><<<grid, hg::kRenThreads, hg::kRenSmemBytes, st>>>(a);
  }
  return hg::check_launch("hg_render_mlp");
}

}  // extern "C"
