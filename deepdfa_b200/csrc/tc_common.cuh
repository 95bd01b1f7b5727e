// Shared device helpers for the tcgen05 kernels: mbarrier, TMA bulk copy, UMMA descriptors, TMEM access,
// the bf16 hi/lo operand split and the "activation image" layout.
//
// Activation image (tcgen05 engine only): an [N,128] fp32 matrix X is also kept as MMA-ready operands:
//   image[tile = node/128][variant v: 0 = hi, 1 = lo][kblock kb: cols 0-63 | 64-127] = one 16 KB chunk,
//   chunk = [128 rows x 64 bf16], rows 128 B apart, 16-byte units XOR-swizzled by (row & 7)
//   (the UMMA canonical K-major SWIZZLE_128B layout; read as "MN-major" it is the [col][node] operand of the
//   weight-gradient GEMM).  hi = bf16(x), lo = bf16(x - hi).  Rows past N are zero.  64 KB per 128-node tile —
//   exactly the bytes of the fp32 matrix — so producer kernels write it instead of / next to fp32 and the GEMM
//   kernels stream it with plain 1-D TMA bulk copies (no in-kernel conversion pass).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace ddfa {
namespace tcc {

constexpr int kD = 128;
constexpr int kTileM = 128;
constexpr int kChunkBytes = 128 * 128;        // one [128 x 64] bf16 chunk
constexpr int kImageTileBytes = 4 * kChunkBytes;  // [hi|lo][kb0|kb1] = 64 KB per 128-node tile

__host__ __device__ __forceinline__ size_t image_bytes(int64_t n) { return (size_t)((n + kTileM - 1) / kTileM) * kImageTileBytes; }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
// true in exactly one lane of a fully converged warp (lets ptxas issue the uniform-datapath tcgen05 / bulk-copy instructions
// straight-line instead of wrapping each one in a loop over "any active lane")
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xffffffff;\n"
      "@px mov.s32 %0, 1;\n"
      "}\n"
      : "+r"(pred));
  return pred != 0;
}
// descriptor of the same matrix `byte_off` further on in shared memory (the start-address field counts 16-byte units)
__device__ __forceinline__ uint64_t desc_advance(uint64_t desc, uint32_t byte_off) { return desc + (uint64_t)(byte_off >> 4); }

// the same with an L2 eviction-priority policy (common.cuh: l2_policy)
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols));
}

// Instruction descriptor, kind::f16: D = f32, A = B = bf16, M = 128; N and operand majors as given.
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn_major = false, bool b_mn_major = false) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
// K-major SWIZZLE_128B shared-memory matrix descriptor: start>>4 | LBO=1 | SBO = 1024 B | version 1 | layout 2
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// MN-major SWIZZLE_128B descriptor: LBO = byte stride between 64-element MN blocks, SBO = 1024 B between 8-row K groups
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with the A operand read from tensor memory: A = [128 lanes = rows of the M dimension][K, two bf16 per 32-bit column]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// registers -> tensor memory: this thread's lane, 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- CTA pairs (cta_group::2): two CTAs of a 2-cluster issue ONE MMA with M = 256 — each CTA's tensor memory holds its own 128
// rows of A and of D, the B operand (N x K) is split by rows of N between the two CTAs' shared memories (same offsets), the leader
// (cluster rank 0) issues, and tcgen05.commit multicasts the completion to mbarriers of both CTAs ------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {      // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {      // arrive on an mbarrier of another CTA of the cluster
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// cta-scope wait, bounded like the one below (used by every role of a pair-form kernel)
__device__ __forceinline__ void mbar_wait_trap(uint32_t bar, uint32_t parity) {
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
// wait with cluster-scope acquire (the arrivals come from the peer CTA); bounded: a protocol error traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t cols) {      // one warp of EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols));
}
// D[tmem, 256 rows over the pair] (+)= A[tmem of each CTA] * B[shared memory of both CTAs]; issued by one thread of the leader CTA
__device__ __forceinline__ void umma_f16_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0)
      : "memory");
}
// completion of all MMAs issued so far -> one arrival on the mbarrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(cta_mask)
               : "memory");
}
// instruction descriptor with M = 256 (cta_group::2)
__host__ __device__ constexpr uint32_t make_idesc_m256(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((256u >> 4) << 24);
}

// byte offset of element (row, k) inside a [rows x 64] bf16 K-major SWIZZLE_128B chunk
__host__ __device__ __forceinline__ uint32_t sw128_offset(int row, int k) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2);
}
// byte offset of element (node, col) variant v inside an activation image
__host__ __device__ __forceinline__ size_t image_offset(int64_t node, int col, int v) {
  return (size_t)(node / kTileM) * kImageTileBytes + (size_t)((v * 2 + (col >> 6)) * kChunkBytes) +
         sw128_offset((int)(node % kTileM), col & 63);
}

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// 4 consecutive fp32 -> packed 4 x bf16 hi and 4 x bf16 lo (8 bytes each).  Two values per conversion instruction
// (cvt.rn.bf16x2.f32 = F2FP.BF16.F32.PACK_AB, which also does the packing) instead of eight scalar F2F on the quarter-rate
// conversion pipe plus shifts and ORs; the values are those of split_bf16 (both round to nearest even).
__device__ __forceinline__ uint32_t bf16x2_bits(float lo_half, float hi_half) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo_half, hi_half);      // .x -> bits 0-15, .y -> bits 16-31
  return *reinterpret_cast<const uint32_t *>(&v);
}
__device__ __forceinline__ void split4(const float4 &x, uint2 &ph, uint2 &pl) {
  ph.x = bf16x2_bits(x.x, x.y);
  ph.y = bf16x2_bits(x.z, x.w);
  pl.x = bf16x2_bits(x.x - __uint_as_float(ph.x << 16), x.y - __uint_as_float(ph.x & 0xffff0000u));
  pl.y = bf16x2_bits(x.z - __uint_as_float(ph.y << 16), x.w - __uint_as_float(ph.y & 0xffff0000u));
}
// 8 consecutive fp32 (one 16-byte bf16 unit) -> hi / lo uint4
__device__ __forceinline__ void split8(const float (&x)[8], uint4 &ph, uint4 &pl) {
  uint2 h0, l0, h1, l1;
  split4(make_float4(x[0], x[1], x[2], x[3]), h0, l0);
  split4(make_float4(x[4], x[5], x[6], x[7]), h1, l1);
  ph = make_uint4(h0.x, h0.y, h1.x, h1.y);
  pl = make_uint4(l0.x, l0.y, l1.x, l1.y);
}

// ---- saved gate values of one element, 64 bits: r, z in [0,1] as 14-bit fixed point, n in [-1,1] as 16-bit fixed point, gh_n as a
// 20-bit float (1 sign, 5 exponent bits with fp16's bias, 14 mantissa bits).  Absolute error <= 3.1e-5 on r, z, 1.6e-5 on n, relative
// 3.1e-5 on gh_n (|gh_n| < 6.1e-5 flushes to 0) — an order below plain fp16 (2.4e-4 / 4.9e-4), which measurably moved the parameter
// gradients (profiles/r03b: 1.6e-5 -> 2e-4 relative), at the same 8 bytes.  Layout: x = r | z << 14 | gh[3:0] << 28 ; y = n | gh[19:4] << 16.
__device__ __forceinline__ uint2 pack_gates(float r, float z, float n, float ghn) {
  // r, z come out of fast_sigmoid (in [0,1]) and n out of fast_tanh (in [-1,1]): no clamping needed.  Rounding to the nearest
  // integer by adding 2^23 (1.5 * 2^23 for the signed value) inside an FMA and reading the low mantissa bits: one FFMA on the
  // main pipe instead of FMUL + F2I (the forward epilogue's conversion / MUFU pipe is its busiest unit), and a single rounding.
  const uint32_t rq = __float_as_uint(fmaf(r, 16383.f, 8388608.f)) & 0x3fffu, zq = __float_as_uint(fmaf(z, 16383.f, 8388608.f)) & 0x3fffu;
  const uint32_t nq = __float_as_uint(fmaf(n, 32767.f, 12582912.f)) & 0xffffu;
  const uint32_t b = __float_as_uint(ghn);
  // round the mantissa to 14 bits (a carry runs into the exponent, as it should), drop the sign, re-bias the exponent 127 -> 15:
  // core = [exponent - 112 | mantissa] ; below 2^-14 -> 0, above fp16's range -> largest value
  const int core = (int)(((b + 0x100u) << 1) >> 10) - (112 << 14);
  uint32_t g = core < (1 << 14) ? 0u : (uint32_t)min(core, (31 << 14) - 1);
  g |= (b >> 12) & 0x80000u;
  return make_uint2(rq | (zq << 14) | (g << 28), nq | ((g >> 4) << 16));
}
__device__ __forceinline__ void unpack_gates(const uint2 &p, float &r, float &z, float &n, float &ghn) {
  r = (float)(p.x & 0x3fffu) * (1.f / 16383.f);
  z = (float)((p.x >> 14) & 0x3fffu) * (1.f / 16383.f);
  // signed 16-bit -> float without I2F: (value + 32768) placed in the mantissa of 2^23, minus (2^23 + 32768) — exact
  n = (__uint_as_float(((p.y & 0xffffu) ^ 0x4b008000u)) - 8421376.f) * (1.f / 32767.f);
  const uint32_t g = ((p.y >> 16) << 4) | (p.x >> 28);
  const uint32_t e = (g >> 14) & 31u;
  ghn = e == 0u ? 0.f : __uint_as_float(((g >> 19) << 31) | ((e + 112u) << 23) | ((g & 0x3fffu) << 9));
}

// Gate math on the MUFU pipe.  __expf() expands to ex2.approx WITHOUT .ftz plus a range fix-up (FSETP, two predicated FMULs) per
// call — three calls per output element in the forward epilogue, which is bound by its instruction count; flushing the (here
// irrelevant) denormal results instead saves 9 instructions per element.  ex2.approx.ftz: 2^-22 relative; rcp.approx.ftz: 1 ulp.
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_sigmoid(float x) { return rcp_ftz(1.f + ex2_ftz(-1.4426950408889634f * x)); }
__device__ __forceinline__ float fast_tanh(float x) {
  const float e = ex2_ftz(2.885390081777927f * x);  // e^(2x): inf for large x -> 1 - 0 = 1 ; 0 for very negative x -> 1 - 2 = -1
  return fmaf(-2.f, rcp_ftz(e + 1.f), 1.f);
}

// ---- pipeline timeline (development aid; ddfa_debug_set key 2 switches it on, ddfa_debug_read fetches it) -----------
// Each translation unit that includes this header gets its own buffer: [CTA][tile][event] SM-clock stamps.
constexpr int kTraceCtas = 148, kTraceTiles = 12, kTraceEvents = 12;
constexpr size_t kTraceWords = (size_t)kTraceCtas * kTraceTiles * kTraceEvents;
static __device__ long long g_trace[kTraceWords];
static __device__ int g_trace_on = 0;
__device__ __forceinline__ void trace_stamp(int on, int tile_i, int ev) {
  if (on && blockIdx.x < kTraceCtas && tile_i < kTraceTiles)
    g_trace[((size_t)blockIdx.x * kTraceTiles + tile_i) * kTraceEvents + ev] = clock64();
}

}  // namespace tcc
}  // namespace ddfa
