// Weight gradient of the discriminator's 3x3 / 1x1 convolutions (autograd through nn.Conv2d in
// lib/discriminators/unet_discriminators.py:21-38), one launch per GROUP of filter taps and per 256-channel chunk
// (as many taps as fit the 512 TMEM columns: ntaps * M-halves * ceil32(Cin chunk) <= 512, so dy is read once per group):
//     dW[co, ci, ky, kx] = sum_{b,h,w} dy[b, co, h, w] * x[b, ci, h + ky - pad, w + kx - pad]
// Same machine as the SPADE weight gradient (csrc/synth_bwd.cu): K = pixels, both operands are K-major as stored
// (NCHW planes are contiguous along W), the operand warps convert rows of 64 pixels into bf16 hi/lo SW128 images --
// the x rows read through the tap's shift with zero padding at the image border -- and the [256 x Cin] fp32
// accumulator stays in TMEM for the CTA's lifetime; per-CTA partials are reduced in fp64 in a fixed order.
#include "common.cuh"
#include "umma.cuh"

namespace hg {

constexpr int kDwThreads = 288;
constexpr uint32_t kDwImg = 256 * 128;
constexpr uint32_t kDwSmemBytes = 4 * kDwImg + 8 * 8 + 16 + 1024;

struct ConvWgradArgs {
  const float* dy;       // [B,Cout,H,W]
  const float* x;        // [B,Cin,H,W]
  float* part_w;         // [grid,256,nq]
  float* part_b;         // [grid,256]
  int B, H, W, Cout, Cin;
  int co0, nco;          // rows of dy handled by this launch (nco <= 256)
  int ci0, nci, nq;      // rows of x (nci valid, nq = nci rounded up to 32, <= 256)
  int ntaps;             // taps of this launch; tap t reads x at (h + oy[t], w + ox[t])
  int oy[9], ox[9];
  // whole-layer mode (hg_conv2d_wgrad_layer): blockIdx.y enumerates (256-row chunk of dy, 256-row chunk of x, group of `per`
  // taps) and the fields above are derived from it; every item owns `item_stride` floats of part_w (gridDim.x partials)
  int layer, ksize, per;
  long item_stride;
};

// item index -> chunk / tap-group geometry, shared by the GEMM kernel and its reduction
struct ConvWgradItem {
  int co0, nco, ci0, nci, nq, ntaps, t0;
};
__device__ __forceinline__ ConvWgradItem conv_wgrad_item(int idx, int Cout, int Cin, int ksize, int per) {
  const int kk = ksize * ksize;
  const int ngrp = (kk + per - 1) / per, nci_chunks = (Cin + 255) / 256;
  const int grp = idx % ngrp, cii = (idx / ngrp) % nci_chunks, coi = idx / (ngrp * nci_chunks);
  ConvWgradItem it;
  it.co0 = coi * 256;
  it.nco = Cout - it.co0 < 256 ? Cout - it.co0 : 256;
  it.ci0 = cii * 256;
  it.nci = Cin - it.ci0 < 256 ? Cin - it.ci0 : 256;
  it.nq = (it.nci + 31) / 32 * 32;
  it.t0 = grp * per;
  it.ntaps = kk - it.t0 < per ? kk - it.t0 : per;
  return it;
}

enum { DW_AFULL = 0, DW_AEMPTY = 1, DW_BFULL = 2, DW_BEMPTY = 3, DW_DONE = 4 };

template <int kPasses>
__global__ void __launch_bounds__(kDwThreads, 1) conv_wgrad_kernel(ConvWgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* s = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_hi = s;
  uint8_t* a_lo = s + kDwImg;
  uint8_t* b_hi = s + 2 * kDwImg;
  uint8_t* b_lo = s + 3 * kDwImg;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s + 4 * kDwImg);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bars + DW_AFULL, 8);
    mbar_init(bars + DW_AEMPTY, 1);
    mbar_init(bars + DW_BFULL, 8);
    mbar_init(bars + DW_BEMPTY, 1);
    mbar_init(bars + DW_DONE, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (a.layer) {
    const ConvWgradItem it = conv_wgrad_item(blockIdx.y, a.Cout, a.Cin, a.ksize, a.per);
    a.co0 = it.co0; a.nco = it.nco; a.ci0 = it.ci0; a.nci = it.nci; a.nq = it.nq; a.ntaps = it.ntaps;
    const int pad = a.ksize >> 1;
    for (int t = 0; t < it.ntaps; ++t) {
      a.oy[t] = (it.t0 + t) / a.ksize - pad;
      a.ox[t] = (it.t0 + t) % a.ksize - pad;
    }
    a.part_w += static_cast<long>(blockIdx.y) * a.item_stride;
    a.part_b += static_cast<long>(blockIdx.y) * gridDim.x * 256;
  }
  const int HW = a.H * a.W, T = (HW + 127) / 128;
  const int total = a.B * T;
  const int count = (total - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int nmh = a.nco > 128 ? 2 : 1;          // M halves that carry rows
  const int nst_a = (a.nco + 31) >> 5, nst_b = a.nq >> 5;

  if (warp < 8) {
    const int sub = threadIdx.x & 7, rsub = threadIdx.x >> 3;
    float bsum[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) bsum[i] = 0.f;
    uint32_t chunk = 0, bcnt = 0;
    for (int it = 0; it < count; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int b = tile / T, ti = tile - b * T;
      const float* dbase = a.dy + (static_cast<long>(b) * a.Cout + a.co0) * HW;
      const float* xbase = a.x + (static_cast<long>(b) * a.Cin + a.ci0) * HW;
#pragma unroll 1
      for (int kc = 0; kc < 2; ++kc, ++chunk) {
        const int g0 = ti * 128 + kc * 64 + sub * 8;        // first pixel of this thread's 8
        const int nvalid = HW - g0;
        float4 va[16];
#pragma unroll
        for (int st = 0; st < 8; ++st) {
          const int row = st * 32 + rsub;
          if (st < nst_a && row < a.nco && nvalid >= 8) {
            const float4* src = reinterpret_cast<const float4*>(dbase + static_cast<long>(row) * HW + g0);
            va[2 * st] = __ldcs(src);
            va[2 * st + 1] = __ldcs(src + 1);
          } else {
            va[2 * st] = va[2 * st + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (st < nst_a && row < a.nco && nvalid > 0) {      // ragged end of the image
              const float* src = dbase + static_cast<long>(row) * HW + g0;
              float t[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) t[j] = j < nvalid ? src[j] : 0.f;
              va[2 * st] = make_float4(t[0], t[1], t[2], t[3]);
              va[2 * st + 1] = make_float4(t[4], t[5], t[6], t[7]);
            }
          }
        }
        mbar_wait_sleep(bars + DW_AEMPTY, (chunk & 1) ^ 1);
#pragma unroll
        for (int st = 0; st < 8; ++st) {
          if (st >= nmh * 4) break;
          const float y[8] = {va[2 * st].x, va[2 * st].y, va[2 * st].z, va[2 * st].w,
                              va[2 * st + 1].x, va[2 * st + 1].y, va[2 * st + 1].z, va[2 * st + 1].w};
          bsum[st] += ((y[0] + y[1]) + (y[2] + y[3])) + ((y[4] + y[5]) + (y[6] + y[7]));
          store_a8<kPasses == 3>(a_hi, a_lo, st * 32 + rsub, sub * 8, y);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + DW_AFULL);
        // ---- x rows through each tap's shift (zero padding outside the image); re-reads hit L1/L2
        int ph[8], pw[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int g = g0 + j;
          ph[j] = g / a.W;
          pw[j] = g - ph[j] * a.W;
        }
#pragma unroll 1
        for (int t = 0; t < a.ntaps; ++t, ++bcnt) {
          const int oy = a.oy[t], ox = a.ox[t];
          bool ok[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            ok[j] = j < nvalid && ph[j] + oy >= 0 && ph[j] + oy < a.H && pw[j] + ox >= 0 && pw[j] + ox < a.W;
          const int shift = oy * a.W + ox;
          // all loads of this tap first (registers only), so that they overlap the MMAs of the previous tap;
          // the shared image is touched only after the EMPTY wait
          float yq[8][8];
#pragma unroll
          for (int st = 0; st < 8; ++st) {
            const int row = st * 32 + rsub;
            if (st < nst_b && row < a.nci) {
              const float* src = xbase + static_cast<long>(row) * HW + g0 + shift;
#pragma unroll
              for (int j = 0; j < 8; ++j) yq[st][j] = ok[j] ? __ldg(src + j) : 0.f;
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) yq[st][j] = 0.f;
            }
          }
          mbar_wait_sleep(bars + DW_BEMPTY, (bcnt & 1) ^ 1);
#pragma unroll
          for (int st = 0; st < 8; ++st) {
            if (st >= nst_b) break;
            store_a8<kPasses == 3>(b_hi, b_lo, st * 32 + rsub, sub * 8, yq[st]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(bars + DW_BFULL);
        }
      }
    }
#pragma unroll
    for (int st = 0; st < 8; ++st) {
      float v = bsum[st];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      if (sub == 0) a.part_b[static_cast<long>(blockIdx.x) * 256 + st * 32 + rsub] = v;
    }
  } else {      // MMA issuer: the warp walks the loops, one elected lane issues (umma.cuh: elect_one_sync)
    const bool leader = elect_one_sync();
    const uint32_t idesc = umma_idesc_bf16(128, a.nq);
    uint32_t chunk = 0, bcnt = 0;
    for (int it = 0; it < count; ++it)
      for (int kc = 0; kc < 2; ++kc, ++chunk) {
        mbar_wait_sleep(bars + DW_AFULL, chunk & 1);
        for (int t = 0; t < a.ntaps; ++t, ++bcnt) {
          mbar_wait_sleep(bars + DW_BFULL, bcnt & 1);
          tc_fence_after();
          for (int mh = 0; mh < nmh; ++mh) {
            const uint32_t d = tmem + (t * nmh + mh) * a.nq;
            const uint32_t ah = smem_u32(a_hi) + mh * (kDwImg / 2), al = smem_u32(a_lo) + mh * (kDwImg / 2);
            umma_k64_if(leader, d, ah, smem_u32(b_hi), idesc, chunk > 0);
            if (kPasses == 3) {
              umma_k64_if(leader, d, al, smem_u32(b_hi), idesc, true);
              umma_k64_if(leader, d, ah, smem_u32(b_lo), idesc, true);
            }
          }
          umma_commit_if(leader, bars + DW_BEMPTY);
        }
        umma_commit_if(leader, bars + DW_AEMPTY);
      }
    umma_commit_if(leader, bars + DW_DONE);
  }
  if (warp < 4) {
    float* dst0 = a.part_w + static_cast<long>(blockIdx.x) * a.ntaps * 256 * a.nq;
    if (count > 0) {
      mbar_wait_long(bars + DW_DONE, 0);
      tc_fence_after();
      for (int t = 0; t < a.ntaps; ++t) {
        float* dst = dst0 + static_cast<long>(t) * 256 * a.nq;
        for (int mh = 0; mh < 2; ++mh) {
          const int co = mh * 128 + warp * 32 + lane;
          for (int cg = 0; cg < (a.nq >> 5); ++cg) {
            uint32_t raw[32];
            if (mh < nmh) {
              tmem_ld32(tmem + (t * nmh + mh) * a.nq + (static_cast<uint32_t>(warp * 32) << 16) + cg * 32, raw);
              tmem_ld_wait();
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) raw[j] = 0u;
            }
            float4* o = reinterpret_cast<float4*>(dst + static_cast<long>(co) * a.nq + cg * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              o[j] = make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]), __uint_as_float(raw[4 * j + 2]),
                                 __uint_as_float(raw[4 * j + 3]));
          }
        }
      }
    } else {
      for (int i = threadIdx.x; i < a.ntaps * 256 * a.nq; i += 128) dst0[i] = 0.f;
    }
  }
  if (count == 0 && warp >= 4 && warp < 8) {
    for (int i = threadIdx.x - 128; i < 256; i += 128) a.part_b[static_cast<long>(blockIdx.x) * 256 + i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tmem);
}

__global__ void conv_wgrad_reduce_kernel(const float* __restrict__ part_w, const float* __restrict__ part_b, int nparts,
                                         int nw, float* __restrict__ dw, float* __restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nw) {
    double acc = 0.0;
    for (int p = 0; p < nparts; ++p) acc += static_cast<double>(part_w[static_cast<long>(p) * nw + i]);
    dw[i] = static_cast<float>(acc);
  }
  if (db && i < 256) {
    double acc = 0.0;
    for (int p = 0; p < nparts; ++p) acc += static_cast<double>(part_b[static_cast<long>(p) * 256 + i]);
    db[i] = static_cast<float>(acc);
  }
}

// whole-layer reduction: item partials -> dW [Cout, Cin, k, k] (and dbias from the items of the first x chunk / tap group), fp64,
// partials in ascending CTA order (deterministic)
__global__ void conv_wgrad_reduce_layer_kernel(const float* __restrict__ part_w, const float* __restrict__ part_b, int nparts,
                                               long item_stride, int Cout, int Cin, int ksize, int per, float* __restrict__ dW,
                                               float* __restrict__ db) {
  const ConvWgradItem it = conv_wgrad_item(blockIdx.y, Cout, Cin, ksize, per);
  const int nw = it.ntaps * 256 * it.nq;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float* pw = part_w + static_cast<long>(blockIdx.y) * item_stride;
  if (i < nw) {
    const int t = i / (256 * it.nq), r = (i / it.nq) % 256, c = i % it.nq;
    if (r < it.nco && c < it.nci) {
      double acc = 0.0;
      for (int p = 0; p < nparts; ++p) acc += static_cast<double>(pw[static_cast<long>(p) * nw + i]);
      dW[(static_cast<long>(it.co0 + r) * Cin + it.ci0 + c) * (ksize * ksize) + it.t0 + t] = static_cast<float>(acc);
    }
  }
  if (db && it.ci0 == 0 && it.t0 == 0 && i < it.nco) {
    const float* pb = part_b + static_cast<long>(blockIdx.y) * nparts * 256;
    double acc = 0.0;
    for (int p = 0; p < nparts; ++p) acc += static_cast<double>(pb[static_cast<long>(p) * 256 + i]);
    db[it.co0 + i] = static_cast<float>(acc);
  }
}

}  // namespace hg

extern "C" {

// worst case per CTA: ntaps * 256 * nq floats with ntaps * nq <= 512 (one M half), plus the bias partial
size_t hg_conv2d_wgrad_workspace_bytes(void) {
  return static_cast<size_t>(hg::num_sms()) * (512 * 256 + 256) * sizeof(float);
}

int hg_conv2d_wgrad_taps(const float* dy, const float* x, float* dw, float* dbias, void* workspace, int B, int H, int W,
                         int Cout, int Cin, int co0, int nco, int ci0, int nci, int ntaps, const int* oy, const int* ox,
                         int passes, void* stream) {
  HG_REQUIRE(dy && x && dw && workspace && oy && ox, "hg_conv2d_wgrad_taps: null pointer");
  HG_REQUIRE(B > 0 && H > 0 && W > 0 && (H * W) % 4 == 0, "hg_conv2d_wgrad_taps: bad image shape (H*W must be a multiple of 4)");
  HG_REQUIRE(nco >= 1 && nco <= 256 && co0 >= 0 && co0 + nco <= Cout, "hg_conv2d_wgrad_taps: bad output-channel chunk");
  HG_REQUIRE(nci >= 1 && nci <= 256 && ci0 >= 0 && ci0 + nci <= Cin, "hg_conv2d_wgrad_taps: bad input-channel chunk");
  HG_REQUIRE(passes == 1 || passes == 3, "hg_conv2d_wgrad_taps: passes must be 1 or 3");
  HG_REQUIRE(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(workspace)) & 15) == 0,
             "hg_conv2d_wgrad_taps: dy / workspace must be 16-byte aligned");
  const int nq = (nci + 31) / 32 * 32;
  const int nmh = nco > 128 ? 2 : 1;
  HG_REQUIRE(ntaps >= 1 && ntaps <= 9 && ntaps * nmh * nq <= 512,
             "hg_conv2d_wgrad_taps: %d taps x %d M-halves x %d columns do not fit the 512 TMEM columns", ntaps, nmh, nq);
  hg::ConvWgradArgs a{};
  for (int t = 0; t < ntaps; ++t) {
    HG_REQUIRE(oy[t] >= -1 && oy[t] <= 1 && ox[t] >= -1 && ox[t] <= 1, "hg_conv2d_wgrad_taps: tap shift out of range");
    a.oy[t] = oy[t];
    a.ox[t] = ox[t];
  }
  const int T = (H * W + 127) / 128;
  const int tiles = B * T;
  const int grid = tiles < hg::num_sms() ? tiles : hg::num_sms();
  float* part_w = static_cast<float*>(workspace);
  float* part_b = part_w + static_cast<size_t>(hg::num_sms()) * 512 * 256;
  a.dy = dy; a.x = x; a.part_w = part_w; a.part_b = part_b;
  a.B = B; a.H = H; a.W = W; a.Cout = Cout; a.Cin = Cin;
  a.co0 = co0; a.nco = nco; a.ci0 = ci0; a.nci = nci; a.nq = nq; a.ntaps = ntaps;
  auto st = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (passes == 3) {
    e = cudaFuncSetAttribute(hg::conv_wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, hg::kDwSmemBytes);
    if (e == cudaSuccess) hg::conv_wgrad_kernel<3><<<grid, hg::kDwThreads, hg::kDwSmemBytes, st>>>(a);
  } else {
    e = cudaFuncSetAttribute(hg::conv_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, hg::kDwSmemBytes);
    if (e == cudaSuccess) hg::conv_wgrad_kernel<1><<<grid, hg::kDwThreads, hg::kDwSmemBytes, st>>>(a);
  }
  if (e != cudaSuccess) { hg::set_error("hg_conv2d_wgrad_taps: smem opt-in failed: %s", cudaGetErrorString(e)); return 2; }
  int rc = hg::check_launch("hg_conv2d_wgrad_taps");
  if (rc) return rc;
  // dw [ntaps, 256, nq] (rows >= nco and columns >= nci are zero), dbias [256]
  const int nw = ntaps * 256 * nq;
  hg::conv_wgrad_reduce_kernel<<<(nw + 255) / 256, 256, 0, st>>>(part_w, part_b, grid, nw, dw, dbias);
  return hg::check_launch("hg_conv2d_wgrad_taps(reduce)");
}

// ---- one launch per layer: every (dy chunk, x chunk, tap group) as blockIdx.y of the same grid.  The low-resolution layers
// (16^2 .. 64^2 pixels: 2 .. 32 tiles per image) needed 9 .. 36 launches of 16-128 CTAs each with hg_conv2d_wgrad_taps.
static void wgrad_layer_geometry(int B, int H, int W, int Cout, int Cin, int ksize, int* per, int* nitems, int* gx, long* item_stride) {
  const int kk = ksize * ksize;
  const int nmh = Cout > 128 ? 2 : 1;
  const int nq_max = ((Cin < 256 ? Cin : 256) + 31) / 32 * 32;
  *per = 512 / (nmh * nq_max) < 1 ? 1 : 512 / (nmh * nq_max);
  if (*per > kk) *per = kk;
  const int ngrp = (kk + *per - 1) / *per;
  *nitems = ((Cout + 255) / 256) * ((Cin + 255) / 256) * ngrp;
  const int tiles = B * ((H * W + 127) / 128);
  // CTAs in flight: one per SM for a single item (every extra CTA is one more partial to drain and reduce), ~4 waves when the
  // grid is many small items of different cost
  int g = *nitems == 1 ? hg::num_sms() : 4 * hg::num_sms() / *nitems;
  if (g < 1) g = 1;
  *gx = tiles < g ? tiles : g;
  *item_stride = static_cast<long>(*gx) * *per * 256 * nq_max;
}

size_t hg_conv2d_wgrad_layer_workspace_bytes(int B, int H, int W, int Cout, int Cin, int ksize) {
  int per, nitems, gx;
  long stride;
  wgrad_layer_geometry(B, H, W, Cout, Cin, ksize, &per, &nitems, &gx, &stride);
  return (static_cast<size_t>(nitems) * stride + static_cast<size_t>(nitems) * gx * 256) * sizeof(float);
}

int hg_conv2d_wgrad_layer(const float* dy, const float* x, float* dW, float* dbias, void* workspace, size_t workspace_bytes, int B,
                          int H, int W, int Cout, int Cin, int ksize, int passes, void* stream) {
  HG_REQUIRE(dy && x && dW && workspace, "hg_conv2d_wgrad_layer: null pointer");
  HG_REQUIRE(ksize == 1 || ksize == 3, "hg_conv2d_wgrad_layer: kernel size must be 1 or 3");
  HG_REQUIRE(B > 0 && H > 0 && W > 0 && (H * W) % 4 == 0, "hg_conv2d_wgrad_layer: bad image shape (H*W must be a multiple of 4)");
  HG_REQUIRE(Cout > 0 && Cin > 0, "hg_conv2d_wgrad_layer: bad channel counts");
  HG_REQUIRE(passes == 1 || passes == 3, "hg_conv2d_wgrad_layer: passes must be 1 or 3");
  HG_REQUIRE(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(workspace)) & 15) == 0,
             "hg_conv2d_wgrad_layer: dy / workspace must be 16-byte aligned");
  int per, nitems, gx;
  long stride;
  wgrad_layer_geometry(B, H, W, Cout, Cin, ksize, &per, &nitems, &gx, &stride);
  HG_REQUIRE(workspace_bytes >= hg_conv2d_wgrad_layer_workspace_bytes(B, H, W, Cout, Cin, ksize),
             "hg_conv2d_wgrad_layer: workspace too small (%zu bytes)", workspace_bytes);
  HG_REQUIRE(nitems <= 65535, "hg_conv2d_wgrad_layer: too many work items");
  hg::ConvWgradArgs a{};
  float* part_w = static_cast<float*>(workspace);
  float* part_b = part_w + static_cast<size_t>(nitems) * stride;
  a.dy = dy; a.x = x; a.part_w = part_w; a.part_b = part_b;
  a.B = B; a.H = H; a.W = W; a.Cout = Cout; a.Cin = Cin;
  a.layer = 1; a.ksize = ksize; a.per = per; a.item_stride = stride;
  auto st = static_cast<cudaStream_t>(stream);
  const dim3 grid(gx, nitems);
  cudaError_t e;
  if (passes == 3) {
    e = cudaFuncSetAttribute(hg::conv_wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, hg::kDwSmemBytes);
    if (e == cudaSuccess) hg::conv_wgrad_kernel<3><<<grid, hg::kDwThreads, hg::kDwSmemBytes, st>>>(a);
  } else {
    e = cudaFuncSetAttribute(hg::conv_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, hg::kDwSmemBytes);
    if (e == cudaSuccess) hg::conv_wgrad_kernel<1><<<grid, hg::kDwThreads, hg::kDwSmemBytes, st>>>(a);
  }
  if (e != cudaSuccess) { hg::set_error("hg_conv2d_wgrad_layer: smem opt-in failed: %s", cudaGetErrorString(e)); return 2; }
  int rc = hg::check_launch("hg_conv2d_wgrad_layer");
  if (rc) return rc;
  const int nw_max = per * 256 * 256;
  hg::conv_wgrad_reduce_layer_kernel<<<dim3((nw_max + 255) / 256, nitems), 256, 0, st>>>(part_w, part_b, gx, stride, Cout, Cin, ksize,
                                                                                          per, dW, dbias);
  return hg::check_launch("hg_conv2d_wgrad_layer(reduce)");
}

}  // extern "C"
