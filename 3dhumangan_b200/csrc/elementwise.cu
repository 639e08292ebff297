// bias_act and upfirdn2d: the two StyleGAN3 native ops the reference ships as CUDA plugins
// (lib/components/ops/bias_act.cu:24-165, lib/components/ops/upfirdn2d.cu:29-375), rebuilt as
// vectorised HBM-streaming kernels for sm_100a.  Both are bandwidth-bound (<= 10 FLOP/B).
#include "common.cuh"

namespace hg {

// activation ids follow bias_act.cpp / bias_act.py:22-32: 1 linear, 2 relu, 3 lrelu, 4 tanh, 5 sigmoid,
// 6 elu, 7 selu, 8 softplus, 9 swish
__device__ __forceinline__ float act_apply(float x, int act, float alpha) {
  switch (act) {
    case 2: return x > 0.f ? x : 0.f;
    case 3: return x > 0.f ? x : x * alpha;
    case 4: return tanhf(x);
    case 5: return 1.f / (1.f + expf(-x));
    case 6: return x > 0.f ? x : expm1f(x);
    case 7: return 1.0507009873554805f * (x > 0.f ? x : 1.6732632423543772f * expm1f(x));
    case 8: return x > 20.f ? x : log1pf(expf(x));
    case 9: return x / (1.f + expf(-x));
    default: return x;
  }
}

__global__ void bias_act_kernel(const float* __restrict__ x, const float* __restrict__ b, float* __restrict__ y,
                                long n, int stepB, int sizeB, int act, float alpha, float gain, float clamp) {
  const long i4 = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  float v[4];
  const bool vec = (i4 + 4 <= n);
  if (vec) {
    const float4 t = *reinterpret_cast<const float4*>(x + i4);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    for (int j = 0; j < 4; ++j) v[j] = (i4 + j < n) ? x[i4 + j] : 0.f;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float t = v[j];
    if (b) t += b[((i4 + j) / stepB) % sizeB];
    t = act_apply(t, act, alpha) * gain;
    if (clamp >= 0.f) t = fminf(fmaxf(t, -clamp), clamp);
    v[j] = t;
  }
  if (vec) {
    *reinterpret_cast<float4*>(y + i4) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    for (int j = 0; j < 4; ++j)
      if (i4 + j < n) y[i4 + j] = v[j];
  }
}


// First / second derivative of act at the pre-activation t, written in terms of the saved forward OUTPUT
// (yy = y / gain) wherever the function allows it, so that backward never needs the forward input
// (bias_act.py:22-32 `ref='y'`; swish is the one activation that needs t itself).  order 1: act'(t), 2: act''(t).
__device__ __forceinline__ float act_derivative(float t, float yy, int act, float alpha, int order) {
  const float kS = 1.0507009873554805f, kSA = 1.0507009873554805f * 1.6732632423543772f;
  if (order == 1) {
    switch (act) {
      case 2: return yy > 0.f ? 1.f : 0.f;
      case 3: return yy > 0.f ? 1.f : alpha;
      case 4: return 1.f - yy * yy;
      case 5: return yy * (1.f - yy);
      case 6: return yy >= 0.f ? 1.f : yy + 1.f;
      case 7: return yy >= 0.f ? kS : yy + kSA;
      case 8: return 1.f - expf(-yy);
      case 9: { const float s = 1.f / (1.f + expf(-t)); return s * (1.f + t * (1.f - s)); }
      default: return 1.f;
    }
  }
  switch (act) {
    case 4: return (1.f - yy * yy) * (-2.f * yy);
    case 5: return yy * (1.f - yy) * (1.f - 2.f * yy);
    case 6: return yy >= 0.f ? 0.f : yy + 1.f;
    case 7: return yy >= 0.f ? 0.f : yy + kSA;
    case 8: { const float c = expf(-yy); return c * (1.f - c); }
    case 9: { const float s = 1.f / (1.f + expf(-t)); const float q = s * (1.f - s); return 2.f * q + t * q * (1.f - 2.f * s); }
    default: return 0.f;
  }
}

// out = g * gain * act^(order)(xref + b) * dy, zero where the forward output was clamped.
__global__ void bias_act_grad_kernel(const float* __restrict__ g, const float* __restrict__ b,
                                     const float* __restrict__ xref, const float* __restrict__ yref,
                                     const float* __restrict__ dy, float* __restrict__ out, long n, int stepB, int sizeB,
                                     int order, int act, float alpha, float gain, float clamp) {
  const long stride = static_cast<long>(gridDim.x) * blockDim.x;
  const float inv_gain = gain != 0.f ? 1.f / gain : 0.f;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float t = xref ? xref[i] : 0.f;
    if (b) t += b[(i / stepB) % sizeB];
    float y = yref ? yref[i] : 0.f;
    if (act == 9) y = act_apply(t, 9, alpha) * gain;       // swish keeps x, not y: rebuild y for the clamp mask
    float v = g[i] * gain * act_derivative(t, y * inv_gain, act, alpha, order);
    if (dy) v *= dy[i];
    if (clamp >= 0.f && !(y > -clamp && y < clamp)) v = 0.f;
    out[i] = v;
  }
}

// out[n,c,oy,ox] = sum_{ky,kx} xup[oy*downy + ky - pady0, ox*downx + kx - padx0] * g[ky,kx]
// where xup is x with (up-1) zeros inserted and g is the (optionally pre-flipped) filter.
__global__ void upfirdn2d_kernel(const float* __restrict__ x, const float* __restrict__ f, float* __restrict__ y,
                                 int NC, int inH, int inW, int outH, int outW, int fH, int fW, int upx, int upy,
                                 int downx, int downy, int padx0, int pady0, int flip, float gain) {
  extern __shared__ float sf[];
  for (int i = threadIdx.x; i < fH * fW; i += blockDim.x) {
    // conv2d is a cross-correlation: the reference flips the filter unless flip_filter (upfirdn2d.py:200-203)
    const int ky = i / fW, kx = i % fW;
    sf[i] = (flip ? f[i] : f[(fH - 1 - ky) * fW + (fW - 1 - kx)]) * gain;
  }
  __syncthreads();
  const long total = static_cast<long>(NC) * outH * outW;
  for (long o = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; o < total;
       o += static_cast<long>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(o % outW);
    const int oy = static_cast<int>((o / outW) % outH);
    const long nc = o / (static_cast<long>(outW) * outH);
    const float* xp = x + nc * inH * inW;
    float acc = 0.f;
    for (int ky = 0; ky < fH; ++ky) {
      const int uy = oy * downy + ky - pady0;
      if (uy < 0 || uy % upy != 0) continue;
      const int iy = uy / upy;
      if (iy >= inH) continue;
      for (int kx = 0; kx < fW; ++kx) {
        const int ux = ox * downx + kx - padx0;
        if (ux < 0 || ux % upx != 0) continue;
        const int ix = ux / upx;
        if (ix >= inW) continue;
        acc = fmaf(xp[static_cast<long>(iy) * inW + ix], sf[ky * fW + kx], acc);
      }
    }
    y[o] = acc;
  }
}

}  // namespace hg

extern "C" {

int hg_bias_act(const float* x, const float* b, float* y, long n, int stepB, int sizeB, int act, float alpha,
                float gain, float clamp, void* stream) {
  HG_REQUIRE(x && y, "hg_bias_act: null pointer");
  HG_REQUIRE(act >= 1 && act <= 9, "hg_bias_act: unknown activation id %d", act);
  HG_REQUIRE(!b || (stepB > 0 && sizeB > 0), "hg_bias_act: bad bias geometry");
  HG_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
             "hg_bias_act: x / y must be 16-byte aligned");
  if (n <= 0) return 0;
  const long threads = (n + 3) / 4;
  hg::bias_act_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, b, y, n, stepB, sizeB, act, alpha, gain, clamp);
  return hg::check_launch("hg_bias_act");
}

int hg_bias_act_grad(const float* g, const float* b, const float* xref, const float* yref, const float* dy, float* out,
                     long n, int stepB, int sizeB, int order, int act, float alpha, float gain, float clamp,
                     void* stream) {
  HG_REQUIRE(g && out, "hg_bias_act_grad: null pointer");
  HG_REQUIRE(act >= 1 && act <= 9, "hg_bias_act_grad: unknown activation id %d", act);
  HG_REQUIRE(order == 1 || order == 2, "hg_bias_act_grad: derivative order must be 1 or 2 (got %d)", order);
  HG_REQUIRE(!b || (stepB > 0 && sizeB > 0), "hg_bias_act_grad: bad bias geometry");
  HG_REQUIRE(act == 1 || (act == 9 ? xref != nullptr : yref != nullptr),
             "hg_bias_act_grad: activation %d needs its saved %s", act, act == 9 ? "input (xref)" : "output (yref)");
  HG_REQUIRE(clamp < 0.f || act == 9 || yref, "hg_bias_act_grad: clamp needs the saved output (yref)");
  if (n <= 0) return 0;
  long blocks = (n + 255) / 256;
  const long cap = static_cast<long>(hg::num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  hg::bias_act_grad_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      g, b, xref, yref, dy, out, n, stepB, sizeB, order, act, alpha, gain, clamp);
  return hg::check_launch("hg_bias_act_grad");
}

int hg_upfirdn2d(const float* x, const float* f, float* y, int NC, int inH, int inW, int outH, int outW, int fH, int fW,
                 int upx, int upy, int downx, int downy, int padx0, int pady0, int flip_filter, float gain,
                 void* stream) {
  HG_REQUIRE(x && f && y, "hg_upfirdn2d: null pointer");
  HG_REQUIRE(upx >= 1 && upy >= 1 && downx >= 1 && downy >= 1, "hg_upfirdn2d: up/down factors must be >= 1");
  HG_REQUIRE(fH >= 1 && fW >= 1 && fH * fW <= 4096, "hg_upfirdn2d: filter too large");
  if (NC <= 0 || outH <= 0 || outW <= 0) return 0;
  const long total = static_cast<long>(NC) * outH * outW;
  long blocks = (total + 255) / 256;
  const long cap = static_cast<long>(hg::num_sms()) * 32;
  if (blocks > cap) blocks = cap;
  hg::upfirdn2d_kernel<<<static_cast<unsigned>(blocks), 256, fH * fW * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      x, f, y, NC, inH, inW, outH, outW, fH, fW, upx, upy, downx, downy, padx0, pady0, flip_filter, gain);
  return hg::check_launch("hg_upfirdn2d");
}

}  // extern "C"
