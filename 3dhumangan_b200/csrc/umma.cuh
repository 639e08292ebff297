// sm_100a primitives shared by every tensor-core kernel in this library: mbarrier, bulk-copy
// (TMA engine, 1-D), tcgen05 MMA / TMEM, and the 128B-swizzled K-major operand layout.
//
// Operand layout (both A and B): "K-major, SWIZZLE_128B" canonical UMMA layout.  A tile of
// R rows x 64 bf16 (= 128 B per row) is stored as R/8 groups of 8 rows; a group is 1024 B; inside
// a group row r (0..7) occupies 128 B and its 16-byte chunk c (0..7) sits at chunk slot (c ^ r).
//     byte(r, k) = (r/8)*1024 + (r%8)*128 + (((k/8) ^ (r%8)) * 16) + (k%8)*2          (k < 64)
// Matrix descriptor: start>>4, LBO=1 (ignored for swizzled K-major), SBO=1024>>4, version=1,
// layout=SWIZZLE_128B(2).  One tcgen05.mma consumes K=16 (32 B): advance the start address by 32 B.
// Tiles must be 1024-byte aligned (base_offset = 0).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hg {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// latency-critical waits (row warps, MMA issuer): plain try_wait loop
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// row-warp waits that are expected to block for a while: a short sleep between polls frees issue slots for the
// other team's warp on the same scheduler (spin loops were 42 % of the executed warp-instructions)
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(32);
}
// producer threads run far ahead of their consumers: back off between polls so that their spin loops do
// not steal issue slots from the row warps on the same scheduler (they were ~50 % of all issued
// warp-instructions, profiles/r1_spade_const_v3_ncu_summary.md)
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(128);
}

// once-per-kernel waits (a drain team that idles until the CTA's last MMA has completed): poll every microsecond
__device__ __forceinline__ void mbar_wait_long(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(1000);
}

// ---------------------------------------------------------------- explicit shared-space accesses
// (pointers that went through integer alignment arithmetic compile to GENERIC LD/ST, which cost extra
//  latency on the hot operand paths; these keep them LDS/STS)
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// Operand-tile store.  Volatile (it has side effects) but WITHOUT a "memory" clobber: the clobber serialised
// every 8-element group behind the previous group's store (table LDS -> math -> STS -> next table LDS ...).
// Ordering against the consumers is provided by fence_proxy_async_smem() (which does clobber memory).
__device__ __forceinline__ void sts_b32x4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d));
}
// Makes the compiler forget what it knows about a register: loads whose address derives from it cannot be
// hoisted above this point (used after a barrier / table refresh in front of NON-volatile table loads).
__device__ __forceinline__ void opaque(uint32_t& r) { asm volatile("" : "+r"(r)); }
// 8 consecutive fp32 entries of a read-mostly shared-memory TABLE (two LDS.128; a warp-wide broadcast LDS.32
// would cost one wavefront per value).  Non-volatile on purpose: the scheduler may batch / hoist them freely;
// callers pass an address made `opaque` after the last point at which the table may have changed.
__device__ __forceinline__ void lds8(uint32_t a, float (&o)[8]) {
  float4 x, y;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(a));
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(y.x), "=f"(y.y), "=f"(y.z), "=f"(y.w) : "r"(a + 16));
  o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w;
  o[4] = y.x; o[5] = y.y; o[6] = y.z; o[7] = y.w;
}

// ---------------------------------------------------------------- proxies / fences
// generic-proxy writes (st.shared) -> visible to the async proxy (tcgen05.mma / bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------- bulk copy global -> shared (TMA engine, 1-D)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------- TMEM
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(kCols) : "memory");
}
// 32 lanes x 32 consecutive columns: thread i of the warp receives lane (base_lane + i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- MMA
// K-major SW128 matrix descriptor for a tile whose 8-row groups are 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);  // start address
  d |= static_cast<uint64_t>(1) << 16;                    // LBO (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;            // SBO: 8 rows * 128 B
  d |= static_cast<uint64_t>(1) << 46;                    // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                    // SWIZZLE_128B
  return d;
}
// Instruction descriptor: kind::f16, BF16 x BF16 -> F32, A and B K-major, dense.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N) {
  return (1u << 4)            // D format F32
         | (1u << 7)          // A format BF16
         | (1u << 10)         // B format BF16
         | ((N >> 3) << 17)   // N / 8
         | ((M >> 4) << 24);  // M / 16
}
// D[tmem] (+)= A[smem] * B[smem]^T   (single thread)
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One leader lane of a CONVERGED warp (the same lane every call).  Loops that issue tcgen05.mma belong in convergent code with
// only the issue itself under `if (leader)`: inside a divergent `if (lane == 0)` region every descriptor lives in vector
// registers and each MMA pays ~5 R2UR + predicate shuffling (measured 17.5 instructions per MMA in the haloed convolution).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// mbarrier arrives once every MMA issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// Predicated forms for issue loops that the WHOLE warp walks (see elect_one_sync): no branch around the instruction.
__device__ __forceinline__ void umma_bf16_if(bool leader, uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(static_cast<uint32_t>(leader))
      : "memory");
}
__device__ __forceinline__ void umma_commit_if(bool leader, uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar)), "r"(static_cast<uint32_t>(leader))
      : "memory");
}
__device__ __forceinline__ void umma_k64_if(bool leader, uint32_t tmem_d, uint32_t a_tile, uint32_t b_tile, uint32_t idesc,
                                            bool accumulate) {
  const uint64_t da = umma_desc_sw128(a_tile);
  const uint64_t db = umma_desc_sw128(b_tile);
#pragma unroll
  for (uint32_t k = 0; k < 4; ++k) umma_bf16_if(leader, tmem_d, da + 2 * k, db + 2 * k, idesc, (accumulate || k > 0) ? 1u : 0u);
}

// One K-chunk of 64: four K=16 MMAs.  a_tile / b_tile: shared addresses of [rows x 64] SW128 tiles.
__device__ __forceinline__ void umma_k64(uint32_t tmem_d, uint32_t a_tile, uint32_t b_tile, uint32_t idesc,
                                         bool accumulate) {
  const uint64_t da = umma_desc_sw128(a_tile);
  const uint64_t db = umma_desc_sw128(b_tile);
#pragma unroll
  for (uint32_t k = 0; k < 4; ++k) {
    // +32 B per K=16 step inside the 128 B swizzle atom (address field is in 16 B units)
    umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (accumulate || k > 0) ? 1u : 0u);
  }
}

// ---------------------------------------------------------------- operand packing
// byte offset of element (row, k) inside a [rows x 64] bf16 SW128 tile
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t k) {
  return (row >> 3) * 1024u + (row & 7u) * 128u + ((((k >> 3) ^ row) & 7u) << 4) + ((k & 7u) << 1);
}

// split two fp32 into (hi, lo) bf16x2 pairs: x ~= hi + lo with |lo| <= 2^-9 |hi|.  The residual x - float(hi) is one packed
// FFMA2 (hi * -1 + x: a single rounding, the same value as the subtraction), the unpack a shift and a mask: 5 instructions per
// pair instead of 6 -- the operand producers of every kernel run this once per element pair and are issue / power bound.
__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  const uint32_t hb = *reinterpret_cast<uint32_t*>(&h);
  const float2 hf = make_float2(__uint_as_float(hb << 16), __uint_as_float(hb & 0xffff0000u));
  const float2 r = __ffma2_rn(hf, make_float2(-1.f, -1.f), make_float2(x0, x1));
  __nv_bfloat162 l = __floats2bfloat162_rn(r.x, r.y);
  hi = hb;
  lo = *reinterpret_cast<uint32_t*>(&l);
}

// y[j] = LeakyReLU_slope(x[j] * g1[j] + g0[j]) for 8 values on packed fp32 (FFMA2 / FMUL2 + FMNMX): max(v, slope * v) is the
// LeakyReLU for 0 <= slope <= 1 (0.2, the identity 1 and ReLU 0 are the slopes this library uses), bit-identical to the
// compare-and-select form for every finite and infinite v (and NaN stays NaN).
__device__ __forceinline__ void affine_lrelu8(const float* __restrict__ x, const float (&g1)[8], const float (&g0)[8], float slope,
                                              float (&y)[8]) {
  const float2 sl = make_float2(slope, slope);
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    const float2 v = __ffma2_rn(make_float2(x[j], x[j + 1]), make_float2(g1[j], g1[j + 1]), make_float2(g0[j], g0[j + 1]));
    const float2 s = __fmul2_rn(v, sl);
    y[j] = fmaxf(v.x, s.x);
    y[j + 1] = fmaxf(v.y, s.y);
  }
}

// Write 8 consecutive-k fp32 values (k0 % 8 == 0) of one row into the hi (and lo) operand tiles.
template <bool kSplit>
__device__ __forceinline__ void store_a8(uint8_t* tile_hi, uint8_t* tile_lo, uint32_t row, uint32_t k0,
                                         const float (&x)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_bf16x2(x[2 * i], x[2 * i + 1], h[i], l[i]);
  const uint32_t off = sw128_offset(row, k0);
  sts_b32x4(smem_u32(tile_hi) + off, h[0], h[1], h[2], h[3]);
  if (kSplit) sts_b32x4(smem_u32(tile_lo) + off, l[0], l[1], l[2], l[3]);
}

}  // namespace hg
