// Backward of one const-style SPADE half-block (forward: csrc/synth.cu; reference: autograd through
// SPADE2d.forward lib/components/map3d_layers.py:176-190 and SPADEBlock.forward :218-238).
//
// With the folded forward  pre = x*g1[b,c] + g0[b,c],  y = lrelu_0.2(pre),  out = W y + bias (+ skip) the
// gradients split into three streaming kernels over the tile-blocked activations [B, T, C=256, 128]:
//
//   hg_spade_bwd_dgrad   (csrc/synth.cu, tcgen05)   dpre = (W^T dout) * lrelu'(pre);  S1[b,c] = sum dpre,
//                                                   S2[b,c] = sum dpre*x
//   hg_spade_bwd_wgrad   (here, tcgen05)            dW[co,ci] = sum_{b,p} dout[b,co,p] * y[b,ci,p]  (y recomputed),
//                                                   dbias[co] = sum dout
//   hg_spade_bwd_combine (here, streaming)          dL/dx = dpre*g1[b,c] + a[c] + k[c]*x  (+ skip gradient)
//                                                   (+ W_rgb^T drgb), and the ToRGB weight gradient
//
// a[c], k[c] carry the gradient that reaches x through the batch statistics (d/dx of sum x and sum x^2); they
// and every other [B,C]/[C]-sized quantity are computed on the host side from S1, S2 (modules/synthesis_bwd.py).
//
// wgrad layout.  Both operands of dW = dout . y^T are K-major in the blocked layout as stored (K = pixels, 128
// contiguous per channel row), so the operand warps only convert rows (coalesced 256 B row segments -> bf16 hi/lo
// SW128 images); per-row constants (g1, g0) instead of per-column ones.  The [256 x 256] fp32 accumulator fills
// the whole TMEM (two M=128 halves x 256 columns) for the CTA's lifetime and is written once, as a per-CTA
// partial, then reduced deterministically by `wgrad_reduce_kernel`.
#include "common.cuh"
#include "umma.cuh"

namespace hg {

constexpr int kWC = 256;
constexpr int kWgThreads = 288;                 // warps 0-7: operand rows, warp 8: MMA issue
constexpr uint32_t kWgImg = 256 * 128;          // [256 rows x 64 px] bf16 = 32 KB
constexpr uint32_t kWgSmemBytes = 4 * kWgImg + 2 * kWC * 4 + 8 * 8 + 16 + 1024;

struct WgradArgs {
  const float* dout;     // [B,T,C,128]
  const float* x;        // [B or 1,T,C,128]
  long x_bstride;
  const float* mod;      // [B,2,C] g1, g0
  float* part_w;         // [grid, C, C]
  float* part_b;         // [grid, C]
  int B, HW;
};

enum { WG_FULL = 0, WG_EMPTY = 1, WG_DONE = 2 };

template <int kPasses>
__global__ void __launch_bounds__(kWgThreads, 1) spade_wgrad_kernel(WgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* s = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_hi = s;
  uint8_t* a_lo = s + kWgImg;
  uint8_t* b_hi = s + 2 * kWgImg;
  uint8_t* b_lo = s + 3 * kWgImg;
  float* tab_g1 = reinterpret_cast<float*>(s + 4 * kWgImg);
  float* tab_g0 = tab_g1 + kWC;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tab_g0 + kWC);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bars + WG_FULL, 8);
    mbar_init(bars + WG_EMPTY, 1);
    mbar_init(bars + WG_DONE, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int T = (a.HW + 127) / 128;
  const int total = a.B * T;
  const int count = (total - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp < 8) {
    const int sub = threadIdx.x & 7;          // which 8-pixel group of the 64-pixel chunk
    const int rsub = threadIdx.x >> 3;        // 0..31: row within a 32-row step
    float bsum[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) bsum[i] = 0.f;
    int cur_b = -1;
    uint32_t chunk = 0;
    for (int it = 0; it < count; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int b = tile / T, ti = tile - b * T;
      if (b != cur_b) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        tab_g1[threadIdx.x] = a.mod[(static_cast<long>(b) * 2 + 0) * kWC + threadIdx.x];
        tab_g0[threadIdx.x] = a.mod[(static_cast<long>(b) * 2 + 1) * kWC + threadIdx.x];
        asm volatile("bar.sync 1, 256;" ::: "memory");
        cur_b = b;
      }
      const float* dbase = a.dout + (static_cast<long>(b) * T + ti) * kWC * 128;
      const float* xbase = a.x + static_cast<long>(b) * a.x_bstride + static_cast<long>(ti) * kWC * 128;
#pragma unroll 1
      for (int kc = 0; kc < 2; ++kc, ++chunk) {
        const int p0 = kc * 64 + sub * 8;                 // first pixel (within the tile) of this thread's 8
        const int nvalid = a.HW - (ti * 128 + p0);        // pixels of the image left from p0 on (may be <= 0)
        // ---- dout rows -> A image.  The loads are issued before the EMPTY wait: they only fill registers.
        float4 va[16];
#pragma unroll
        for (int st = 0; st < 8; ++st) {
          const float4* src = reinterpret_cast<const float4*>(dbase + (st * 32 + rsub) * 128 + p0);
          va[2 * st] = __ldcs(src);
          va[2 * st + 1] = __ldcs(src + 1);
        }
        mbar_wait_sleep(bars + WG_EMPTY, (chunk & 1) ^ 1);
#pragma unroll
        for (int st = 0; st < 8; ++st) {
          float y[8] = {va[2 * st].x, va[2 * st].y, va[2 * st].z, va[2 * st].w,
                        va[2 * st + 1].x, va[2 * st + 1].y, va[2 * st + 1].z, va[2 * st + 1].w};
          if (nvalid < 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = j < nvalid ? y[j] : 0.f;
          }
          bsum[st] += ((y[0] + y[1]) + (y[2] + y[3])) + ((y[4] + y[5]) + (y[6] + y[7]));
          store_a8<kPasses == 3>(a_hi, a_lo, st * 32 + rsub, sub * 8, y);
        }
        // ---- x rows -> y = lrelu(x*g1 + g0) -> B image
#pragma unroll
        for (int st = 0; st < 8; ++st) {
          const float4* src = reinterpret_cast<const float4*>(xbase + (st * 32 + rsub) * 128 + p0);
          va[2 * st] = __ldcs(src);
          va[2 * st + 1] = __ldcs(src + 1);
        }
#pragma unroll
        for (int st = 0; st < 8; ++st) {
          const int row = st * 32 + rsub;
          const float g1 = tab_g1[row], g0 = tab_g0[row];
          float y[8] = {va[2 * st].x, va[2 * st].y, va[2 * st].z, va[2 * st].w,
                        va[2 * st + 1].x, va[2 * st + 1].y, va[2 * st + 1].z, va[2 * st + 1].w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float pre = fmaf(y[j], g1, g0);
            y[j] = j < nvalid ? (pre > 0.f ? pre : 0.2f * pre) : 0.f;
          }
          store_a8<kPasses == 3>(b_hi, b_lo, row, sub * 8, y);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + WG_FULL);
      }
    }
    // bias-gradient partials: row (st*32 + rsub) is shared by the 8 `sub` lanes
#pragma unroll
    for (int st = 0; st < 8; ++st) {
      float v = bsum[st];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      if (sub == 0) a.part_b[static_cast<long>(blockIdx.x) * kWC + st * 32 + rsub] = v;
    }
  } else if (lane == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, 256);
    uint32_t chunk = 0;
    for (int it = 0; it < count; ++it)
      for (int kc = 0; kc < 2; ++kc, ++chunk) {
        mbar_wait_sleep(bars + WG_FULL, chunk & 1);
        tc_fence_after();
#pragma unroll
        for (uint32_t mh = 0; mh < 2; ++mh) {
          const uint32_t d = tmem + mh * 256;
          const uint32_t ah = smem_u32(a_hi) + mh * (kWgImg / 2), al = smem_u32(a_lo) + mh * (kWgImg / 2);
          umma_k64(d, ah, smem_u32(b_hi), idesc, chunk > 0);
          if (kPasses == 3) {
            umma_k64(d, al, smem_u32(b_hi), idesc, true);
            umma_k64(d, ah, smem_u32(b_lo), idesc, true);
          }
        }
        umma_commit(bars + WG_EMPTY);
      }
    umma_commit(bars + WG_DONE);
  }
  // ---- drain: warps 0-3 own TMEM lanes 32w..32w+31 (co within the half), 2 x 256 columns (ci)
  if (warp < 4) {
    float* dst = a.part_w + static_cast<long>(blockIdx.x) * kWC * kWC;
    if (count > 0) {
      mbar_wait_sleep(bars + WG_DONE, 0);
      tc_fence_after();
      for (int mh = 0; mh < 2; ++mh) {
        const int co = mh * 128 + warp * 32 + lane;
        for (int cg = 0; cg < 8; ++cg) {
          uint32_t raw[32];
          tmem_ld32(tmem + mh * 256 + (static_cast<uint32_t>(warp * 32) << 16) + cg * 32, raw);
          tmem_ld_wait();
          float4* o = reinterpret_cast<float4*>(dst + static_cast<long>(co) * kWC + cg * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            o[j] = make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]), __uint_as_float(raw[4 * j + 2]),
                               __uint_as_float(raw[4 * j + 3]));
        }
      }
    } else {
      for (int i = threadIdx.x; i < kWC * kWC; i += 128) dst[i] = 0.f;
    }
  }
  if (count == 0 && warp >= 4 && warp < 8) {     // an idle CTA still owns a (zero) bias partial
    for (int i = threadIdx.x - 128; i < kWC; i += 128) a.part_b[static_cast<long>(blockIdx.x) * kWC + i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tmem);
}

// dW[i] = sum over CTAs of part[cta][i] (fp64 accumulation, fixed order -> deterministic), likewise the bias.
__global__ void wgrad_reduce_kernel(const float* __restrict__ part_w, const float* __restrict__ part_b, int nparts,
                                    float* __restrict__ dw, float* __restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kWC * kWC) {
    double acc = 0.0;
    for (int p = 0; p < nparts; ++p) acc += static_cast<double>(part_w[static_cast<long>(p) * kWC * kWC + i]);
    dw[i] = static_cast<float>(acc);
  }
  if (db && i < kWC) {
    double acc = 0.0;
    for (int p = 0; p < nparts; ++p) acc += static_cast<double>(part_b[static_cast<long>(p) * kWC + i]);
    db[i] = static_cast<float>(acc);
  }
}

// ------------------------------------------------------------------------------------------
// combine: the gradient that reaches the INPUT x of a half-block (= the output of the previous one)
//   dx[b,c,p] = dpre[b,c,p]*g1[b,c] + a[c] + k[c]*x[b,c,p]  (+ dskip[b,c,p])  (+ sum_j W_rgb[j,c]*drgb[b,j,p])
// plus, when x is also the input of a ToRGB layer, dW_rgb[j,c] = sum_{b,p} drgb[b,j,p]*x[b,c,p].
// Pure streaming: warp w owns channels w, w+8, ...; a lane owns 4 consecutive pixels of the 128-pixel tile.
// ------------------------------------------------------------------------------------------
struct CombineArgs {
  const float* dpre;     // [B,T,C,128] or null (then only the skip / rgb terms)
  const float* x;        // [B or 1,T,C,128] (needed when k or rgb_w is given)
  long x_bstride;
  const float* g1;       // [B,2,C] (row 0 used) or null
  const float* ak;       // [2,C]: a, k or null
  const float* dskip;    // [B,T,C,128] or null
  const float* drgb;     // [B,3,HW] or null
  const float* rgb_w;    // [3,C]
  float* dx;             // [B,T,C,128]
  double* dwrgb;         // [3,C] accumulated, or null
  int B, HW;
};

__global__ void __launch_bounds__(256) spade_combine_kernel(CombineArgs a) {
  __shared__ float s_wrgb[3 * kWC];
  __shared__ float s_acc[3 * kWC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 3 * kWC; i += blockDim.x) {
    s_wrgb[i] = a.drgb ? a.rgb_w[i] : 0.f;
    s_acc[i] = 0.f;
  }
  __syncthreads();
  const int T = (a.HW + 127) / 128;
  const int total = a.B * T;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int b = tile / T, ti = tile - b * T;
    const int p = ti * 128 + lane * 4;
    const long off = (static_cast<long>(b) * T + ti) * kWC * 128 + lane * 4;
    const long xoff = static_cast<long>(b) * a.x_bstride + static_cast<long>(ti) * kWC * 128 + lane * 4;
    float4 r[3];
    if (a.drgb) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float* src = a.drgb + (static_cast<long>(b) * 3 + j) * a.HW + p;
        r[j] = p + 3 < a.HW ? *reinterpret_cast<const float4*>(src)
                            : make_float4(p < a.HW ? src[0] : 0.f, p + 1 < a.HW ? src[1] : 0.f, p + 2 < a.HW ? src[2] : 0.f, 0.f);
      }
    }
    const float m0 = p < a.HW ? 1.f : 0.f, m1 = p + 1 < a.HW ? 1.f : 0.f, m2 = p + 2 < a.HW ? 1.f : 0.f,
                m3 = p + 3 < a.HW ? 1.f : 0.f;
#pragma unroll 4
    for (int c = warp; c < kWC; c += 8) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.ak || a.drgb) xv = __ldcs(reinterpret_cast<const float4*>(a.x + xoff + c * 128));
      if (a.dpre) {
        const float4 d = __ldcs(reinterpret_cast<const float4*>(a.dpre + off + c * 128));
        const float g = a.g1[static_cast<long>(b) * 2 * kWC + c];
        v = make_float4(d.x * g, d.y * g, d.z * g, d.w * g);
      }
      if (a.ak) {
        const float aa = a.ak[c], kk = a.ak[kWC + c];
        v.x += fmaf(kk, xv.x, aa); v.y += fmaf(kk, xv.y, aa); v.z += fmaf(kk, xv.z, aa); v.w += fmaf(kk, xv.w, aa);
      }
      if (a.dskip) {
        const float4 d = __ldcs(reinterpret_cast<const float4*>(a.dskip + off + c * 128));
        v.x += d.x; v.y += d.y; v.z += d.z; v.w += d.w;
      }
      if (a.drgb) {
        float t[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float w = s_wrgb[j * kWC + c];
          v.x = fmaf(w, r[j].x, v.x); v.y = fmaf(w, r[j].y, v.y); v.z = fmaf(w, r[j].z, v.z); v.w = fmaf(w, r[j].w, v.w);
          t[j] = (r[j].x * xv.x * m0 + r[j].y * xv.y * m1) + (r[j].z * xv.z * m2 + r[j].w * xv.w * m3);
        }
        if (a.dwrgb) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            t[0] += __shfl_xor_sync(0xffffffffu, t[0], o);
            t[1] += __shfl_xor_sync(0xffffffffu, t[1], o);
            t[2] += __shfl_xor_sync(0xffffffffu, t[2], o);
          }
          if (lane == 0) {   // channel c belongs to this warp alone: no atomics
            s_acc[c] += t[0];
            s_acc[kWC + c] += t[1];
            s_acc[2 * kWC + c] += t[2];
          }
        }
      }
      v.x *= m0; v.y *= m1; v.z *= m2; v.w *= m3;
      __stcs(reinterpret_cast<float4*>(a.dx + off + c * 128), v);
    }
  }
  __syncthreads();
  if (a.dwrgb)
    for (int i = threadIdx.x; i < 3 * kWC; i += blockDim.x) atomicAdd(a.dwrgb + i, static_cast<double>(s_acc[i]));
}


// ------------------------------------------------------------------------------------------
// synthesis input backward: x0[c,p] = sin(w[c,0]*i(p) + w[c,1]*j(p) + b[c]) is shared by the batch, so
//   darg[c,p] = cos(arg) * sum_b dx[b,c,p];  dw[c,0] = sum_p darg*i,  dw[c,1] = sum_p darg*j,  db[c] = sum_p darg
// (autograd through SynthesisInput.forward, map3d_layers.py:260-275).  One block per channel.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) synth_input_bwd_kernel(const float* __restrict__ dx, const float* __restrict__ w,
                                                              const float* __restrict__ bias, const float* __restrict__ ic,
                                                              const float* __restrict__ jc, int B, int Hg, int Wg,
                                                              float* __restrict__ dw, float* __restrict__ db) {
  const int c = blockIdx.x;
  const int HW = Hg * Wg, T = (HW + 127) / 128;
  const float w0 = w[c * 2 + 0], w1 = w[c * 2 + 1], bb = bias[c];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    const float iv = ic[p / Wg], jv = jc[p % Wg];
    float g = 0.f;
    for (int b = 0; b < B; ++b) g += dx[((static_cast<long>(b) * T + (p >> 7)) * kWC + c) * 128 + (p & 127)];
    const float d = g * cosf(fmaf(w1, jv, fmaf(w0, iv, bb)));
    a0 += static_cast<double>(d * iv);
    a1 += static_cast<double>(d * jv);
    a2 += static_cast<double>(d);
  }
  __shared__ double red[3][8];
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a0; red[1][threadIdx.x >> 5] = a1; red[2][threadIdx.x >> 5] = a2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0, t1 = 0, t2 = 0;
    for (int i = 0; i < 8; ++i) { t0 += red[0][i]; t1 += red[1][i]; t2 += red[2][i]; }
    dw[c * 2 + 0] = static_cast<float>(t0);
    dw[c * 2 + 1] = static_cast<float>(t1);
    db[c] = static_cast<float>(t2);
  }
}

}  // namespace hg

extern "C" {

size_t hg_spade_bwd_wgrad_workspace_bytes(void) {
  return static_cast<size_t>(hg::num_sms()) * (hg::kWC * hg::kWC + hg::kWC) * sizeof(float);
}

int hg_spade_bwd_wgrad(const float* dout, const float* x, long x_bstride, const float* mod, float* dw, float* dbias,
                       void* workspace, int B, int C, int Hg, int Wg, int passes, void* stream) {
  HG_REQUIRE(C == hg::kWC, "hg_spade_bwd_wgrad: only %d channels are supported (got %d)", hg::kWC, C);
  HG_REQUIRE(dout && x && mod && dw && workspace, "hg_spade_bwd_wgrad: null pointer");
  HG_REQUIRE(passes == 1 || passes == 3, "hg_spade_bwd_wgrad: passes must be 1 or 3");
  HG_REQUIRE(B > 0 && Hg > 0 && Wg > 0, "hg_spade_bwd_wgrad: bad shape");
  HG_REQUIRE(((reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(workspace)) & 15) == 0,
             "hg_spade_bwd_wgrad: tensors must be 16-byte aligned");
  const int T = (Hg * Wg + 127) / 128;
  const int tiles = B * T;
  const int grid = tiles < hg::num_sms() ? tiles : hg::num_sms();
  float* part_w = static_cast<float*>(workspace);
  float* part_b = part_w + static_cast<size_t>(hg::num_sms()) * hg::kWC * hg::kWC;
  hg::WgradArgs a{dout, x, x_bstride, mod, part_w, part_b, B, Hg * Wg};
  auto st = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (passes == 3) {
    e = cudaFuncSetAttribute(hg::spade_wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, hg::kWgSmemBytes);
    if (e == cudaSuccess) hg::spade_wgrad_kernel<3><<<grid, hg::kWgThreads, hg::kWgSmemBytes, st>>>(a);
  } else {
    e = cudaFuncSetAttribute(hg::spade_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, hg::kWgSmemBytes);
    if (e == cudaSuccess) hg::spade_wgrad_kernel<1><<<grid, hg::kWgThreads, hg::kWgSmemBytes, st>>>(a);
  }
  if (e != cudaSuccess) { hg::set_error("hg_spade_bwd_wgrad: smem opt-in failed: %s", cudaGetErrorString(e)); return 2; }
  int rc = hg::check_launch("hg_spade_bwd_wgrad");
  if (rc) return rc;
  hg::wgrad_reduce_kernel<<<(hg::kWC * hg::kWC + 255) / 256, 256, 0, st>>>(part_w, part_b, grid, dw, dbias);
  return hg::check_launch("hg_spade_bwd_wgrad(reduce)");
}

int hg_spade_bwd_combine(const float* dpre, const float* x, long x_bstride, const float* g1, const float* ak,
                         const float* dskip, const float* drgb, const float* rgb_w, float* dx, double* dwrgb, int B, int C,
                         int Hg, int Wg, void* stream) {
  HG_REQUIRE(C == hg::kWC, "hg_spade_bwd_combine: only %d channels are supported (got %d)", hg::kWC, C);
  HG_REQUIRE(dx, "hg_spade_bwd_combine: null output");
  HG_REQUIRE(!dpre || g1, "hg_spade_bwd_combine: dpre needs its g1 table");
  HG_REQUIRE(!(ak || drgb) || x, "hg_spade_bwd_combine: x is needed for the statistics / ToRGB terms");
  HG_REQUIRE(!drgb || rgb_w, "hg_spade_bwd_combine: drgb needs rgb_w");
  HG_REQUIRE(!drgb || ((Hg * Wg) % 4 == 0 && (reinterpret_cast<uintptr_t>(drgb) & 15) == 0),
             "hg_spade_bwd_combine: drgb must be 16-byte aligned with H*W a multiple of 4");
  HG_REQUIRE(B > 0 && Hg > 0 && Wg > 0, "hg_spade_bwd_combine: bad shape");
  const int T = (Hg * Wg + 127) / 128;
  const int tiles = B * T;
  int grid = hg::num_sms() * 4;
  if (grid > tiles) grid = tiles;
  hg::CombineArgs a{dpre, x, x_bstride, g1, ak, dskip, drgb, rgb_w, dx, dwrgb, B, Hg * Wg};
  hg::spade_combine_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return hg::check_launch("hg_spade_bwd_combine");
}

int hg_synth_input_bwd(const float* dx, const float* w, const float* bias, const float* ic, const float* jc, int B, int C,
                       int Hg, int Wg, float* dw, float* db, void* stream) {
  HG_REQUIRE(C == hg::kWC, "hg_synth_input_bwd: only %d channels are supported (got %d)", hg::kWC, C);
  HG_REQUIRE(dx && w && bias && ic && jc && dw && db, "hg_synth_input_bwd: null pointer");
  HG_REQUIRE(B > 0 && Hg > 0 && Wg > 0, "hg_synth_input_bwd: bad shape");
  hg::synth_input_bwd_kernel<<<C, 256, 0, static_cast<cudaStream_t>(stream)>>>(dx, w, bias, ic, jc, B, Hg, Wg, dw, db);
  return hg::check_launch("hg_synth_input_bwd");
}

}  // extern "C"
