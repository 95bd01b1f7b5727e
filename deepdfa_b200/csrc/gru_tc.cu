// tcgen05 engine for K4 (one GRU propagation step, D == 128) — Blackwell tensor cores with TMEM
// accumulators, fp32-grade accuracy through bf16x3 split operands.
//
//   acc_r   = s W'_r^T + h Whh_r^T      acc_z = s W'_z^T + h Whh_z^T        (K = 128 + 128)
//   acc_gin = s W'_n^T                  acc_ghn = h Whh_n^T                 (K = 128 each)
//   epilogue: r,z = sigmoid(acc + indeg*b' + b_ih + b_hh), n = tanh(gin + r*ghn), h' = n + z (h - n)
//
// Precision: every fp32 operand x is split x = hi + lo with hi = bf16(x), lo = bf16(x - hi); the
// product a*w is accumulated as a_hi*w_hi + a_lo*w_hi + a_hi*w_lo in fp32 (TMEM), dropping only the
// a_lo*w_lo term (~2^-16 relative).  Single-pass bf16/TF32 misses the 1e-3 logit bound at trained
// weight scales (hard part 1).
//
// CTA = one tile of 128 nodes, 10 warps, warp-specialised:
//   warp 0      : TMEM allocator + weight producer.  Streams 24 pre-swizzled 16 KB weight chunks
//                 (UMMA K-major SWIZZLE_128B smem images, built once per forward by
//                 gru_tc_pack_kernel) with cp.async.bulk (TMA 1-D bulk copy) into a 5-stage ring.
//   warp 1      : MMA issuer (one elected lane): 144 x tcgen05.mma.cta_group::1.kind::f16
//                 (M=128, N=128, K=16), accumulators r|z|gin|ghn = 4 x 128 TMEM columns.
//   warps 2..9  : (1) load the s / h node tiles (coalesced 128-bit loads), split to bf16 hi/lo and
//                 store them as swizzled K-major A operands; (2) epilogue: tcgen05.ld the
//                 accumulators (thread = node row), gate math, write h' (and the saved gates).
// Synchronisation is mbarrier-only between roles (full/empty ring, A-ready, accumulators-ready).
#include <cuda_bf16.h>

#include "common.cuh"

namespace ddfa {
namespace tc {

constexpr int kD = 128;
constexpr int kTileM = 128;
constexpr int kChunkBytes = 128 * 128;     // 128 weight rows x 64 bf16 (one 128 B swizzle row each)
constexpr int kNumChunks = 24;             // (s|h) x (kblock 0|1) x (hi|lo) x (r|z|n)
constexpr int kStages = 5;
constexpr int kATileBytes = kTileM * 128;  // one K-block (64 bf16) of one A variant
constexpr int kSmemA = 8 * kATileBytes;    // [part s|h][variant hi|lo][kblock]  = 128 KB
constexpr int kSmemB = kStages * kChunkBytes;
constexpr int kBiasFloats = 7 * kD;        // b_ih_r+b_hh_r | b_ih_z+b_hh_z | b_ih_n | b_hh_n | b'_r | b'_z | b'_n
constexpr int kOffB = kSmemA;
constexpr int kOffBias = kOffB + kSmemB;
constexpr int kOffBar = kOffBias + kBiasFloats * 4;
constexpr int kNumBars = 2 * kStages + 3;  // full[5], empty[5], a_ready[2], acc_ready
constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16;
constexpr int kSmemAlloc = kSmemBytes + 1024;  // slack for the 1024 B alignment of the swizzled tiles
constexpr int kThreads = 320;
constexpr int kLoaderThreads = 256;
constexpr size_t kPackedBytes = (size_t)kNumChunks * kChunkBytes + kBiasFloats * 4;

// Instruction descriptor, kind::f16: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), both K-major,
// N>>3 at bit 17, M>>4 at bit 24.
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) = 1 | SBO>>4 [32,46) = 1024>>4 | version [46,48) = 1 | layout [61,64) = 2
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// byte offset of element (row, k) inside a [rows x 64] bf16 K-major SWIZZLE_128B tile
__host__ __device__ __forceinline__ uint32_t sw128_offset(int row, int k) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2);
}

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// chunk c = ((p*2 + kb)*2 + v)*3 + g   with p: 0 = w_fold (s part), 1 = w_hh (h part); kb: K block;
// v: 0 = hi, 1 = lo; g: gate (r,z,n).  One thread packs 8 consecutive k of one weight row.
__global__ void __launch_bounds__(256) gru_tc_pack_kernel(const float *__restrict__ w_fold, const float *__restrict__ w_hh,
                                                          const float *__restrict__ b_fold, const float *__restrict__ b_ih,
                                                          const float *__restrict__ b_hh, uint8_t *__restrict__ packed) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = 2 * 2 * 3 * 128 * 8;  // (p, kb, g, row, k8): hi and lo are written together
  if (t < total) {
    const int k8 = t & 7;
    const int row = (t >> 3) & 127;
    const int g = (t >> 10) % 3;
    const int pk = t / (3 * 1024);
    const int kb = pk & 1, p = pk >> 1;
    const float *W = (p == 0 ? w_fold : w_hh) + (size_t)(g * 128 + row) * kD + kb * 64 + k8 * 8;
    const float4 a = *reinterpret_cast<const float4 *>(W);
    const float4 b = *reinterpret_cast<const float4 *>(W + 4);
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split_bf16(x[i], hi[i], lo[i]);
    const uint32_t off = sw128_offset(row, k8 * 8);
    const int c_hi = ((p * 2 + kb) * 2 + 0) * 3 + g, c_lo = ((p * 2 + kb) * 2 + 1) * 3 + g;
    *reinterpret_cast<uint4 *>(packed + (size_t)c_hi * kChunkBytes + off) = *reinterpret_cast<const uint4 *>(hi);
    *reinterpret_cast<uint4 *>(packed + (size_t)c_lo * kChunkBytes + off) = *reinterpret_cast<const uint4 *>(lo);
  }
  if (t < kD) {
    float *bias = reinterpret_cast<float *>(packed + (size_t)kNumChunks * kChunkBytes);
    bias[0 * kD + t] = b_ih[t] + b_hh[t];
    bias[1 * kD + t] = b_ih[kD + t] + b_hh[kD + t];
    bias[2 * kD + t] = b_ih[2 * kD + t];
    bias[3 * kD + t] = b_hh[2 * kD + t];
    bias[4 * kD + t] = b_fold[t];
    bias[5 * kD + t] = b_fold[kD + t];
    bias[6 * kD + t] = b_fold[2 * kD + t];
  }
}

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) {
  const float e = __expf(2.f * x);          // inf for large x -> 1 - 0 = 1 ; 0 for very negative x -> -1
  return 1.f - __fdividef(2.f, e + 1.f);
}

__global__ void __launch_bounds__(kThreads, 1) gru_tc_fwd_kernel(const float *__restrict__ s, const float *__restrict__ h,
                                                                 const int32_t *__restrict__ indptr,
                                                                 const uint8_t *__restrict__ packed, int32_t N,
                                                                 float *__restrict__ h_out, float *__restrict__ gates) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + kOffBar;
  auto full_bar = [&](int i) { return bar0 + 8u * i; };
  auto empty_bar = [&](int i) { return bar0 + 8u * (kStages + i); };
  auto aready_bar = [&](int p) { return bar0 + 8u * (2 * kStages + p); };
  const uint32_t acc_bar = bar0 + 8u * (2 * kStages + 2);
  volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(smem + kOffTmemPtr);
  float *bias_s = reinterpret_cast<float *>(smem + kOffBias);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile0 = blockIdx.x * kTileM;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(full_bar(i), 1); mbar_init(empty_bar(i), 1); }
    mbar_init(aready_bar(0), kLoaderThreads);
    mbar_init(aready_bar(1), kLoaderThreads);
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {  // TMEM: all 512 columns (4 accumulators x 128 fp32 columns)
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_ptr_smem)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  {  // biases -> smem
    const float *bias_g = reinterpret_cast<const float *>(packed + (size_t)kNumChunks * kChunkBytes);
    for (int i = threadIdx.x; i < kBiasFloats; i += kThreads) bias_s[i] = bias_g[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== weight producer =====
    if (lane == 0) {
      for (int c = 0; c < kNumChunks; ++c) {
        const int stage = c % kStages, use = c / kStages;
        if (use > 0) mbar_wait(empty_bar(stage), (use - 1) & 1);
        mbar_arrive_expect_tx(full_bar(stage), kChunkBytes);
        bulk_g2s(sbase + kOffB + stage * kChunkBytes, packed + (size_t)c * kChunkBytes, kChunkBytes, full_bar(stage));
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t started = 0;  // bit per accumulator region (r, z, gin, ghn)
      for (int c = 0; c < kNumChunks; ++c) {
        const int g = c % 3, v = (c / 3) & 1, kb = (c / 6) & 1, p = c / 12;
        const int stage = c % kStages, use = c / kStages;
        if (c % 12 == 0) mbar_wait(aready_bar(p), 0);
        mbar_wait(full_bar(stage), use & 1);
        tc_fence_after();
        const int region = (g < 2) ? g : (2 + p);
        const uint32_t d_addr = tmem_base + (uint32_t)region * 128u;
        const uint32_t b_addr = sbase + kOffB + stage * kChunkBytes;
        const int n_av = (v == 0) ? 2 : 1;  // w_hi pairs with a_hi and a_lo; w_lo with a_hi only
        for (int av = 0; av < n_av; ++av) {
          const uint32_t a_addr = sbase + (uint32_t)(((p * 2 + av) * 2 + kb) * kATileBytes);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            umma_f16(d_addr, make_desc(a_addr + k4 * 32), make_desc(b_addr + k4 * 32), kIdesc, (started >> region) & 1u);
            started |= 1u << region;
          }
        }
        umma_commit(empty_bar(stage));  // frees the ring slot once these MMAs have read it
      }
      umma_commit(acc_bar);  // all accumulators complete
    }
  } else {
    // ===== A-operand loaders, then epilogue =====
    const int lw = warp - 2;  // 0..7
    for (int p = 0; p < 2; ++p) {
      const float *src = (p == 0) ? s : h;
      const int kb = lane >> 4;                 // lanes 0-15 -> K block 0, 16-31 -> K block 1
      const int kin = (lane & 15) * 4;          // k inside the block
      uint8_t *a_hi = smem + ((p * 2 + 0) * 2 + kb) * kATileBytes;
      uint8_t *a_lo = smem + ((p * 2 + 1) * 2 + kb) * kATileBytes;
#pragma unroll
      for (int r0 = 0; r0 < 16; r0 += 4) {
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int row = lw * 16 + r0 + j;
          const int node = tile0 + row;
          v[j] = (node < N) ? ldg_nc_f4(src + (size_t)node * kD + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int row = lw * 16 + r0 + j;
          __nv_bfloat16 hi[4], lo[4];
          split_bf16(v[j].x, hi[0], lo[0]); split_bf16(v[j].y, hi[1], lo[1]);
          split_bf16(v[j].z, hi[2], lo[2]); split_bf16(v[j].w, hi[3], lo[3]);
          const uint32_t off = sw128_offset(row, kin);
          uint2 ph, pl;
          ph.x = (uint32_t)__bfloat16_as_ushort(hi[0]) | ((uint32_t)__bfloat16_as_ushort(hi[1]) << 16);
          ph.y = (uint32_t)__bfloat16_as_ushort(hi[2]) | ((uint32_t)__bfloat16_as_ushort(hi[3]) << 16);
          pl.x = (uint32_t)__bfloat16_as_ushort(lo[0]) | ((uint32_t)__bfloat16_as_ushort(lo[1]) << 16);
          pl.y = (uint32_t)__bfloat16_as_ushort(lo[2]) | ((uint32_t)__bfloat16_as_ushort(lo[3]) << 16);
          *reinterpret_cast<uint2 *>(a_hi + off) = ph;
          *reinterpret_cast<uint2 *>(a_lo + off) = pl;
        }
      }
      fence_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
      mbar_arrive(aready_bar(p));
    }

    // ---- epilogue: thread = node row (TMEM lane), 64 columns per warp in 4 chunks of 16 ----
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int chalf = lw >> 2;       // column half
    const int row = q * 32 + lane;
    const int node = tile0 + row;
    const bool valid = node < N;
    const float deg = valid ? (float)(indptr[node + 1] - indptr[node]) : 0.f;
    const size_t plane = (size_t)N * kD;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      const int col0 = chalf * 64 + cc * 16;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col0;
      float ar[16], az[16], agi[16], agh[16];
      tmem_ld16(taddr + 0, ar);
      tmem_ld16(taddr + 128, az);
      tmem_ld16(taddr + 256, agi);
      tmem_ld16(taddr + 384, agh);
      float hv[16];
      if (valid) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t4 = *reinterpret_cast<const float4 *>(h + (size_t)node * kD + col0 + i * 4);
          hv[i * 4 + 0] = t4.x; hv[i * 4 + 1] = t4.y; hv[i * 4 + 2] = t4.z; hv[i * 4 + 3] = t4.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) hv[i] = 0.f;
      }
      tmem_ld_wait();
      float o_r[16], o_z[16], o_n[16], o_g[16], o_h[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = col0 + i;
        const float r = fast_sigmoid(ar[i] + fmaf(deg, bias_s[4 * kD + c], bias_s[0 * kD + c]));
        const float z = fast_sigmoid(az[i] + fmaf(deg, bias_s[5 * kD + c], bias_s[1 * kD + c]));
        const float ghn = agh[i] + bias_s[3 * kD + c];
        const float nn = fast_tanh(agi[i] + fmaf(deg, bias_s[6 * kD + c], bias_s[2 * kD + c]) + r * ghn);
        o_r[i] = r; o_z[i] = z; o_n[i] = nn; o_g[i] = ghn;
        o_h[i] = fmaf(z, hv[i] - nn, nn);
      }
      if (valid) {
        float *dst = h_out + (size_t)node * kD + col0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<float4 *>(dst + i * 4) = make_float4(o_h[i * 4], o_h[i * 4 + 1], o_h[i * 4 + 2], o_h[i * 4 + 3]);
        if (gates) {
          float *gd = gates + (size_t)node * kD + col0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            *reinterpret_cast<float4 *>(gd + i * 4) = make_float4(o_r[i * 4], o_r[i * 4 + 1], o_r[i * 4 + 2], o_r[i * 4 + 3]);
            *reinterpret_cast<float4 *>(gd + plane + i * 4) = make_float4(o_z[i * 4], o_z[i * 4 + 1], o_z[i * 4 + 2], o_z[i * 4 + 3]);
            *reinterpret_cast<float4 *>(gd + 2 * plane + i * 4) = make_float4(o_n[i * 4], o_n[i * 4 + 1], o_n[i * 4 + 2], o_n[i * 4 + 3]);
            *reinterpret_cast<float4 *>(gd + 3 * plane + i * 4) = make_float4(o_g[i * 4], o_g[i * 4 + 1], o_g[i * 4 + 2], o_g[i * 4 + 3]);
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// =================================================================================================
// Backward, part 1 — gate backward fused with the two data-gradient GEMMs (K = 3D):
//     q_r, q_z, q_n, q_nr  (elementwise from dh', h, r, z, n, gh_n)           dgi = [q_r|q_z|q_n]
//     ds = dgi  W'      (= q_r W'_r + q_z W'_z + q_n  W'_n)                   dgh = [q_r|q_z|q_nr]
//     dh = dh' * z + dgh Whh (= q_r Whh_r + q_z Whh_z + q_nr Whh_n)
// Also writes the four q planes (inputs of the weight-gradient kernel) and the bias gradients.
// Same CTA organisation as the forward kernel; the A operand cycles the four q matrices through
// two 64 KB slots (a_full / a_empty mbarriers), accumulators ds | dh = 2 x 128 TMEM columns.
// Weight chunks (B operands, [n = output col][k = gate col] K-major, i.e. the TRANSPOSED weights) are
// pre-packed by gru_tc_pack_bwd_kernel in consumption order:
//   c in [ 0, 8): A = q_r : target ds (W'_r^T)  c<4, target dh (Whh_r^T) c>=4 ; kb = (c>>1)&1 ; v = c&1
//   c in [ 8,16): A = q_z : likewise with gate z
//   c in [16,20): A = q_n : ds (W'_n^T) ;  c in [20,24): A = q_nr : dh (Whh_n^T)
// =================================================================================================
__host__ __device__ __forceinline__ void bwd_chunk_decode(int c, int &m, int &target, int &kb, int &v) {
  if (c < 16) { m = c >> 3; target = (c >> 2) & 1; }
  else        { m = 2 + ((c - 16) >> 2); target = m - 2; }
  kb = (c >> 1) & 1;
  v = c & 1;
}

__global__ void __launch_bounds__(256) gru_tc_pack_bwd_kernel(const float *__restrict__ w_fold, const float *__restrict__ w_hh,
                                                              uint8_t *__restrict__ packed) {
  // one thread = (chunk pair cp (hi+lo), k8, n): 8 consecutive k of output column n
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 12 * 8 * 128) return;
  const int n = t & 127, k8 = (t >> 7) & 7, cp = t >> 10;
  int m, target, kb, v;
  bwd_chunk_decode(cp * 2, m, target, kb, v);
  const int gate = (m < 2) ? m : 2;
  const float *W = (target == 0) ? w_fold : w_hh;
  __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float x = W[(size_t)(gate * 128 + kb * 64 + k8 * 8 + i) * kD + n];   // B[n][k] = W[gate*128 + k][n]
    split_bf16(x, hi[i], lo[i]);
  }
  const uint32_t off = sw128_offset(n, k8 * 8);
  *reinterpret_cast<uint4 *>(packed + (size_t)(cp * 2 + 0) * kChunkBytes + off) = *reinterpret_cast<const uint4 *>(hi);
  *reinterpret_cast<uint4 *>(packed + (size_t)(cp * 2 + 1) * kChunkBytes + off) = *reinterpret_cast<const uint4 *>(lo);
}

constexpr int kBwdSlotBytes = 4 * kATileBytes;        // one q matrix: [hi|lo][kblock] = 64 KB
constexpr int kBwdOffB = 2 * kBwdSlotBytes;           // 128 KB
constexpr int kBwdOffBias = kBwdOffB + kSmemB;
constexpr int kBwdOffBar = kBwdOffBias + kBiasFloats * 4;
constexpr int kBwdNumBars = 2 * kStages + 5;          // full[5], empty[5], a_full[2], a_empty[2], acc
constexpr int kBwdOffTmemPtr = kBwdOffBar + kBwdNumBars * 8;
constexpr int kBwdSmemAlloc = kBwdOffTmemPtr + 16 + 1024;

__device__ __forceinline__ void store_split_row(uint8_t *slot, int kb, uint32_t off, const float4 &x) {
  __nv_bfloat16 hi[4], lo[4];
  split_bf16(x.x, hi[0], lo[0]); split_bf16(x.y, hi[1], lo[1]);
  split_bf16(x.z, hi[2], lo[2]); split_bf16(x.w, hi[3], lo[3]);
  uint2 ph, pl;
  ph.x = (uint32_t)__bfloat16_as_ushort(hi[0]) | ((uint32_t)__bfloat16_as_ushort(hi[1]) << 16);
  ph.y = (uint32_t)__bfloat16_as_ushort(hi[2]) | ((uint32_t)__bfloat16_as_ushort(hi[3]) << 16);
  pl.x = (uint32_t)__bfloat16_as_ushort(lo[0]) | ((uint32_t)__bfloat16_as_ushort(lo[1]) << 16);
  pl.y = (uint32_t)__bfloat16_as_ushort(lo[2]) | ((uint32_t)__bfloat16_as_ushort(lo[3]) << 16);
  *reinterpret_cast<uint2 *>(slot + (0 * 2 + kb) * kATileBytes + off) = ph;   // variant hi
  *reinterpret_cast<uint2 *>(slot + (1 * 2 + kb) * kATileBytes + off) = pl;   // variant lo
}

__global__ void __launch_bounds__(kThreads, 1) gru_tc_dgrad_kernel(
    const float *__restrict__ dh_out, const float *__restrict__ h, const float *__restrict__ gates,
    const int32_t *__restrict__ indptr, const uint8_t *__restrict__ packed, int32_t N, float *__restrict__ ds,
    float *__restrict__ dh, float *__restrict__ q, float *__restrict__ db_fold, float *__restrict__ db_ih,
    float *__restrict__ db_hh) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + kBwdOffBar;
  auto full_bar = [&](int i) { return bar0 + 8u * i; };
  auto empty_bar = [&](int i) { return bar0 + 8u * (kStages + i); };
  auto afull_bar = [&](int s_) { return bar0 + 8u * (2 * kStages + s_); };
  auto aempty_bar = [&](int s_) { return bar0 + 8u * (2 * kStages + 2 + s_); };
  const uint32_t acc_bar = bar0 + 8u * (2 * kStages + 4);
  volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(smem + kBwdOffTmemPtr);
  float *bias_s = reinterpret_cast<float *>(smem + kBwdOffBias);   // 7 x 128 column sums of this tile

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile0 = blockIdx.x * kTileM;
  const size_t plane = (size_t)N * kD;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(full_bar(i), 1); mbar_init(empty_bar(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(afull_bar(i), kLoaderThreads); mbar_init(aempty_bar(i), 1); }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_ptr_smem)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = threadIdx.x; i < kBiasFloats; i += kThreads) bias_s[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      for (int c = 0; c < kNumChunks; ++c) {
        const int stage = c % kStages, use = c / kStages;
        if (use > 0) mbar_wait(empty_bar(stage), (use - 1) & 1);
        mbar_arrive_expect_tx(full_bar(stage), kChunkBytes);
        bulk_g2s(sbase + kBwdOffB + stage * kChunkBytes, packed + (size_t)c * kChunkBytes, kChunkBytes, full_bar(stage));
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t started = 0;  // bit per accumulator (ds, dh)
      for (int c = 0; c < kNumChunks; ++c) {
        int m, target, kb, v;
        bwd_chunk_decode(c, m, target, kb, v);
        const int slot = m & 1;
        const bool first_of_m = (c == 0) || (c == 8) || (c == 16) || (c == 20);
        const bool last_of_m = (c == 7) || (c == 15) || (c == 19) || (c == 23);
        if (first_of_m) mbar_wait(afull_bar(slot), (m >> 1) & 1);
        const int stage = c % kStages, use = c / kStages;
        mbar_wait(full_bar(stage), use & 1);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + (uint32_t)target * 128u;
        const uint32_t b_addr = sbase + kBwdOffB + stage * kChunkBytes;
        const int n_av = (v == 0) ? 2 : 1;
        for (int av = 0; av < n_av; ++av) {
          const uint32_t a_addr = sbase + (uint32_t)(slot * kBwdSlotBytes + (av * 2 + kb) * kATileBytes);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            umma_f16(d_addr, make_desc(a_addr + k4 * 32), make_desc(b_addr + k4 * 32), kIdesc, (started >> target) & 1u);
            started |= 1u << target;
          }
        }
        umma_commit(empty_bar(stage));
        if (last_of_m) umma_commit(aempty_bar(slot));   // this q matrix has been fully consumed
      }
      umma_commit(acc_bar);
    }
  } else {
    const int lw = warp - 2;
    const int kb = lane >> 4, kin = (lane & 15) * 4;
    const int col = lane * 4;
    float4 sum[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) sum[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    // ---- sweep 1: q_r -> slot 0, q_z -> slot 1; all four q planes to global ----
#pragma unroll 2
    for (int r = 0; r < 16; ++r) {
      const int row = lw * 16 + r;
      const int node = tile0 + row;
      float4 qr = make_float4(0.f, 0.f, 0.f, 0.f), qz = qr, qn = qr, qnr = qr;
      if (node < N) {
        const size_t off = (size_t)node * kD + col;
        const float4 d = ldg_nc_f4(dh_out + off);
        const float4 hv = ldg_nc_f4(h + off);
        const float4 rr = ldg_nc_f4(gates + off);
        const float4 zz = ldg_nc_f4(gates + plane + off);
        const float4 nn = ldg_nc_f4(gates + 2 * plane + off);
        const float4 gh = ldg_nc_f4(gates + 3 * plane + off);
        const float deg = (float)(indptr[node + 1] - indptr[node]);
#define BWDQ(f)                                                  \
  {                                                              \
    const float dz_ = d.f * (hv.f - nn.f);                       \
    const float dn_ = d.f * (1.f - zz.f);                        \
    qn.f = dn_ * (1.f - nn.f * nn.f);                            \
    qz.f = dz_ * zz.f * (1.f - zz.f);                            \
    qr.f = qn.f * gh.f * rr.f * (1.f - rr.f);                    \
    qnr.f = qn.f * rr.f;                                         \
  }
        BWDQ(x) BWDQ(y) BWDQ(z) BWDQ(w)
#undef BWDQ
        *reinterpret_cast<float4 *>(q + off) = qr;
        *reinterpret_cast<float4 *>(q + plane + off) = qz;
        *reinterpret_cast<float4 *>(q + 2 * plane + off) = qn;
        *reinterpret_cast<float4 *>(q + 3 * plane + off) = qnr;
        f4_add(sum[0], qr); f4_add(sum[1], qz); f4_add(sum[2], qn); f4_add(sum[3], qnr);
        f4_fma(sum[4], deg, qr); f4_fma(sum[5], deg, qz); f4_fma(sum[6], deg, qn);
      }
      const uint32_t off_s = sw128_offset(row, kin);
      store_split_row(smem + 0 * kBwdSlotBytes, kb, off_s, qr);
      store_split_row(smem + 1 * kBwdSlotBytes, kb, off_s, qz);
    }
    fence_async_smem();
    mbar_arrive(afull_bar(0));
    mbar_arrive(afull_bar(1));
    // column sums of this tile -> shared (8 warps contend per address), later one RED per column per CTA
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      atomicAdd(&bias_s[i * kD + col + 0], sum[i].x); atomicAdd(&bias_s[i * kD + col + 1], sum[i].y);
      atomicAdd(&bias_s[i * kD + col + 2], sum[i].z); atomicAdd(&bias_s[i * kD + col + 3], sum[i].w);
    }
    // ---- sweep 2: q_n -> slot 0, q_nr -> slot 1 (re-read this thread's own q values) ----
#pragma unroll 1
    for (int m = 2; m < 4; ++m) {
      const int slot = m & 1;
      mbar_wait(aempty_bar(slot), 0);
      const float *qp = q + (size_t)m * plane;
#pragma unroll 4
      for (int r = 0; r < 16; ++r) {
        const int row = lw * 16 + r;
        const int node = tile0 + row;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (node < N) x = *reinterpret_cast<const float4 *>(qp + (size_t)node * kD + col);
        store_split_row(smem + slot * kBwdSlotBytes, kb, sw128_offset(row, kin), x);
      }
      fence_async_smem();
      mbar_arrive(afull_bar(slot));
    }
    // bias gradients: the 256 loader threads flush the CTA's column sums
    asm volatile("bar.sync 1, 256;" ::: "memory");
    for (int i = threadIdx.x - 64; i < kBiasFloats; i += kLoaderThreads) {
      const float v_ = bias_s[i];
      const int which = i >> 7, c_ = i & 127;
      // 0:S(q_r) 1:S(q_z) 2:S(q_n) 3:S(q_nr) 4:S(deg q_r) 5:S(deg q_z) 6:S(deg q_n)
      if (which == 0) { atomicAdd(db_ih + c_, v_); atomicAdd(db_hh + c_, v_); }
      else if (which == 1) { atomicAdd(db_ih + kD + c_, v_); atomicAdd(db_hh + kD + c_, v_); }
      else if (which == 2) atomicAdd(db_ih + 2 * kD + c_, v_);
      else if (which == 3) atomicAdd(db_hh + 2 * kD + c_, v_);
      else atomicAdd(db_fold + (which - 4) * kD + c_, v_);
    }
    // ---- epilogue: ds = acc_ds ; dh = acc_dh + dh' * z ----
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const int qd = warp & 3, chalf = lw >> 2;
    const int row = qd * 32 + lane;
    const int node = tile0 + row;
    const bool valid = node < N;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      const int col0 = chalf * 64 + cc * 16;
      const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)col0;
      float a_ds[16], a_dh[16];
      tmem_ld16(taddr + 0, a_ds);
      tmem_ld16(taddr + 128, a_dh);
      float dv[16], zv[16];
      if (valid) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t4 = *reinterpret_cast<const float4 *>(dh_out + (size_t)node * kD + col0 + i * 4);
          const float4 z4 = *reinterpret_cast<const float4 *>(gates + plane + (size_t)node * kD + col0 + i * 4);
          dv[i * 4 + 0] = t4.x; dv[i * 4 + 1] = t4.y; dv[i * 4 + 2] = t4.z; dv[i * 4 + 3] = t4.w;
          zv[i * 4 + 0] = z4.x; zv[i * 4 + 1] = z4.y; zv[i * 4 + 2] = z4.z; zv[i * 4 + 3] = z4.w;
        }
      }
      tmem_ld_wait();
      if (valid) {
        float *pds = ds + (size_t)node * kD + col0;
        float *pdh = dh + (size_t)node * kD + col0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          *reinterpret_cast<float4 *>(pds + i * 4) = make_float4(a_ds[i * 4], a_ds[i * 4 + 1], a_ds[i * 4 + 2], a_ds[i * 4 + 3]);
          *reinterpret_cast<float4 *>(pdh + i * 4) =
              make_float4(fmaf(dv[i * 4], zv[i * 4], a_dh[i * 4]), fmaf(dv[i * 4 + 1], zv[i * 4 + 1], a_dh[i * 4 + 1]),
                          fmaf(dv[i * 4 + 2], zv[i * 4 + 2], a_dh[i * 4 + 2]), fmaf(dv[i * 4 + 3], zv[i * 4 + 3], a_dh[i * 4 + 3]));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
  }
}

// =================================================================================================
// Backward, part 2 — weight gradients (K = nodes):  dW'[3D,D] += dgi^T s ,  dWhh[3D,D] += dgh^T h.
// Both operands are "MN-major" (for one node k the 128 gate / feature columns are contiguous): the
// same swizzled [node][64 cols] tiles the other kernels use, read through MN-major UMMA descriptors
// (LBO = stride between the two 64-column blocks, SBO = stride between 8-node groups).
// grid = (ctas, 2): blockIdx.y = role (0: A in {q_r,q_z,q_n}, B = s -> dW' ; 1: A in {q_r,q_z,q_nr}, B = h
// -> dWhh).  A CTA is persistent over 64-node tiles and keeps its three [128 x 128] fp32 accumulator
// blocks (384 TMEM columns) across all of them; the operands stream through a 6-slot ring (B_t, A_0, A_1,
// A_2 per tile; B_t is released after A_2).  At the end every CTA adds its partial sums with RED.ADD.
// =================================================================================================
constexpr int kWgTileK = 64;                          // nodes per tile
constexpr int kWgSlotBytes = 2 * 2 * kWgTileK * 128;  // [hi|lo][col block][64 nodes x 128 B] = 32 KB
constexpr int kWgSlots = 6;
constexpr int kWgOffBar = kWgSlots * kWgSlotBytes;    // 192 KB
constexpr int kWgNumBars = 2 * kWgSlots + 1;
constexpr int kWgOffTmemPtr = kWgOffBar + kWgNumBars * 8;
constexpr int kWgSmemAlloc = kWgOffTmemPtr + 16 + 1024;
// both operands MN-major: bit 15 (A) and bit 16 (B)
constexpr uint32_t kIdescMN = kIdesc | (1u << 15) | (1u << 16);

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
  // start>>4 | LBO (between 64-element MN blocks) = 8192 B | SBO (between 8-row K groups) = 1024 B | v1 | SW128
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((kWgTileK * 128) >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__global__ void __launch_bounds__(kThreads, 1) gru_tc_wgrad_kernel(const float *__restrict__ q, const float *__restrict__ s,
                                                                   const float *__restrict__ h, int32_t N,
                                                                   float *__restrict__ dw_fold, float *__restrict__ dw_hh) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + kWgOffBar;
  auto full_bar = [&](int i) { return bar0 + 8u * i; };
  auto empty_bar = [&](int i) { return bar0 + 8u * (kWgSlots + i); };
  const uint32_t acc_bar = bar0 + 8u * (2 * kWgSlots);
  volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(smem + kWgOffTmemPtr);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int role = blockIdx.y;
  const int num_tiles = (N + kWgTileK - 1) / kWgTileK;
  const int my_tiles = (num_tiles > (int)blockIdx.x) ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const size_t plane = (size_t)N * kD;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgSlots; ++i) { mbar_init(full_bar(i), kLoaderThreads); mbar_init(empty_bar(i), 1); }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_ptr_smem)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 1) {
    if (lane == 0 && my_tiles > 0) {
      for (int i = 0; i < my_tiles; ++i) {
        const int jb = 4 * i;                       // operand sequence number of B_t
        const int slot_b = jb % kWgSlots;
        mbar_wait(full_bar(slot_b), (jb / kWgSlots) & 1);
        for (int g = 0; g < 3; ++g) {
          const int ja = jb + 1 + g, slot_a = ja % kWgSlots;
          mbar_wait(full_bar(slot_a), (ja / kWgSlots) & 1);
          tc_fence_after();
          const uint32_t a0 = sbase + slot_a * kWgSlotBytes, b0 = sbase + slot_b * kWgSlotBytes;
          const uint32_t d_addr = tmem_base + (uint32_t)g * 128u;
#pragma unroll
          for (int k16 = 0; k16 < kWgTileK / 16; ++k16) {
            const uint32_t koff = (uint32_t)k16 * 2048u;   // 16 nodes = two 8-node groups of 1024 B
            const uint32_t vstride = 2 * kWgTileK * 128;   // hi -> lo variant
            const uint32_t acc = (i > 0 || k16 > 0) ? 1u : 0u;
            umma_f16(d_addr, make_desc_mn(a0 + koff), make_desc_mn(b0 + koff), kIdescMN, acc);                     // a_hi b_hi
            umma_f16(d_addr, make_desc_mn(a0 + vstride + koff), make_desc_mn(b0 + koff), kIdescMN, 1u);            // a_lo b_hi
            umma_f16(d_addr, make_desc_mn(a0 + koff), make_desc_mn(b0 + vstride + koff), kIdescMN, 1u);            // a_hi b_lo
          }
          umma_commit(empty_bar(slot_a));
        }
        umma_commit(empty_bar(slot_b));
      }
      umma_commit(acc_bar);
    }
  } else if (warp >= 2) {
    const int lw = warp - 2;
    const int kb = lane >> 4, kin = (lane & 15) * 4;
    const int col = lane * 4;
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      const int node0 = tile * kWgTileK;
      for (int w = 0; w < 4; ++w) {
        const int j = 4 * i + w, slot = j % kWgSlots;
        if (j >= kWgSlots) mbar_wait(empty_bar(slot), ((j / kWgSlots) - 1) & 1);
        const float *src;
        if (w == 0) src = (role == 0) ? s : h;
        else {
          const int pl = (w == 3) ? (role == 0 ? 2 : 3) : (w - 1);   // q planes: r, z, then n (role 0) or nr (role 1)
          src = q + (size_t)pl * plane;
        }
        uint8_t *dst = smem + slot * kWgSlotBytes;
        float4 v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int node = node0 + lw * 8 + r;
          v[r] = (node < N) ? ldg_nc_f4(src + (size_t)node * kD + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int row = lw * 8 + r;
          __nv_bfloat16 hi[4], lo[4];
          split_bf16(v[r].x, hi[0], lo[0]); split_bf16(v[r].y, hi[1], lo[1]);
          split_bf16(v[r].z, hi[2], lo[2]); split_bf16(v[r].w, hi[3], lo[3]);
          uint2 ph, pl2;
          ph.x = (uint32_t)__bfloat16_as_ushort(hi[0]) | ((uint32_t)__bfloat16_as_ushort(hi[1]) << 16);
          ph.y = (uint32_t)__bfloat16_as_ushort(hi[2]) | ((uint32_t)__bfloat16_as_ushort(hi[3]) << 16);
          pl2.x = (uint32_t)__bfloat16_as_ushort(lo[0]) | ((uint32_t)__bfloat16_as_ushort(lo[1]) << 16);
          pl2.y = (uint32_t)__bfloat16_as_ushort(lo[2]) | ((uint32_t)__bfloat16_as_ushort(lo[3]) << 16);
          const uint32_t off = sw128_offset(row, kin);
          *reinterpret_cast<uint2 *>(dst + (0 * 2 + kb) * (kWgTileK * 128) + off) = ph;
          *reinterpret_cast<uint2 *>(dst + (1 * 2 + kb) * (kWgTileK * 128) + off) = pl2;
        }
        fence_async_smem();
        mbar_arrive(full_bar(slot));
      }
    }
    if (my_tiles > 0) {
      mbar_wait(acc_bar, 0);
      tc_fence_after();
      const int qd = warp & 3, chalf = lw >> 2;
      const int m = qd * 32 + lane;                      // row inside the 128-row gate block
      float *dW = (role == 0) ? dw_fold : dw_hh;
#pragma unroll 1
      for (int g = 0; g < 3; ++g) {
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int col0 = chalf * 64 + cc * 16;
          float a[16];
          tmem_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(g * 128 + col0), a);
          tmem_ld_wait();
          float *dst = dW + (size_t)(g * 128 + m) * kD + col0;
#pragma unroll
          for (int x = 0; x < 16; ++x) atomicAdd(dst + x, a[x]);
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

}  // namespace tc

bool gru_tc_available() { return true; }

size_t gru_tc_workspace_bytes(int32_t N, int32_t D) {
  (void)N;
  return D == tc::kD ? tc::kPackedBytes : 16;
}

int gru_tc_prepare(const float *w_fold, const float *b_fold, const float *b_ih, const float *w_hh, const float *b_hh, int32_t D,
                   void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (D != tc::kD) {
    set_error("tcgen05 engine: D must be 128, got %d", D);
    return DDFA_ERR_UNSUPPORTED;
  }
  if (workspace == nullptr || workspace_bytes < tc::kPackedBytes) {
    set_error("tcgen05 engine: workspace too small (%zu < %zu)", workspace_bytes, tc::kPackedBytes);
    return DDFA_ERR_WORKSPACE;
  }
  const int total = 2 * 2 * 3 * 128 * 8;
  tc::gru_tc_pack_kernel<<<(total + 255) / 256, 256, 0, stream>>>(w_fold, w_hh, b_fold, b_ih, b_hh, static_cast<uint8_t *>(workspace));
  DDFA_CHECK_LAUNCH("gru_tc_pack_kernel");
  return DDFA_OK;
}

// ---- backward host side --------------------------------------------------------------------------
// workspace = [24 x 16 KB packed transposed weights][q planes: 4 x N x 128 fp32]
size_t gru_tc_bwd_workspace_bytes(int32_t N, int32_t D) {
  if (D != tc::kD) return 16;
  return (size_t)tc::kNumChunks * tc::kChunkBytes + (size_t)4 * (size_t)N * tc::kD * sizeof(float);
}

int gru_tc_prepare_bwd(const float *w_fold, const float *w_hh, int32_t D, void *workspace, size_t workspace_bytes,
                       cudaStream_t stream) {
  if (D != tc::kD) {
    set_error("tcgen05 engine: D must be 128, got %d", D);
    return DDFA_ERR_UNSUPPORTED;
  }
  if (workspace == nullptr || workspace_bytes < (size_t)tc::kNumChunks * tc::kChunkBytes) {
    set_error("tcgen05 engine (bwd): workspace too small");
    return DDFA_ERR_WORKSPACE;
  }
  tc::gru_tc_pack_bwd_kernel<<<(12 * 8 * 128 + 255) / 256, 256, 0, stream>>>(w_fold, w_hh, static_cast<uint8_t *>(workspace));
  DDFA_CHECK_LAUNCH("gru_tc_pack_bwd_kernel");
  return DDFA_OK;
}

int gru_tc_step_bwd(const float *dh_out, const float *h, const float *s, const float *gates, const int32_t *indptr, int32_t N,
                    int32_t D, float *ds, float *dh, float *dw_fold, float *db_fold, float *db_ih, float *dw_hh, float *db_hh,
                    void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (D != tc::kD) {
    set_error("tcgen05 engine: D must be 128, got %d", D);
    return DDFA_ERR_UNSUPPORTED;
  }
  if (workspace == nullptr || workspace_bytes < gru_tc_bwd_workspace_bytes(N, D)) {
    set_error("tcgen05 engine (bwd): workspace too small (%zu < %zu)", workspace_bytes, gru_tc_bwd_workspace_bytes(N, D));
    return DDFA_ERR_WORKSPACE;
  }
  uint8_t *packed = static_cast<uint8_t *>(workspace);
  float *q = reinterpret_cast<float *>(packed + (size_t)tc::kNumChunks * tc::kChunkBytes);
  DDFA_CUDA(cudaFuncSetAttribute(tc::gru_tc_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kBwdSmemAlloc));
  DDFA_CUDA(cudaFuncSetAttribute(tc::gru_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kWgSmemAlloc));
  const int tiles = (N + tc::kTileM - 1) / tc::kTileM;
  tc::gru_tc_dgrad_kernel<<<tiles, tc::kThreads, tc::kBwdSmemAlloc, stream>>>(dh_out, h, gates, indptr, packed, N, ds, dh, q, db_fold,
                                                                            db_ih, db_hh);
  DDFA_CHECK_LAUNCH("gru_tc_dgrad_kernel");
  const int wtiles = (N + tc::kWgTileK - 1) / tc::kWgTileK;
  int ctas = kNumSMs / 2;
  if (ctas > wtiles) ctas = wtiles;
  dim3 grid(ctas, 2);
  tc::gru_tc_wgrad_kernel<<<grid, tc::kThreads, tc::kWgSmemAlloc, stream>>>(q, s, h, N, dw_fold, dw_hh);
  DDFA_CHECK_LAUNCH("gru_tc_wgrad_kernel");
  return DDFA_OK;
}

int gru_tc_step_fwd(const float *s, const float *h, const int32_t *indptr, const float *w_fold, const float *b_fold,
                    const float *b_ih, const float *w_hh, const float *b_hh, int32_t N, int32_t D, float *h_out,
                    float *save_gates, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  (void)w_fold; (void)b_fold; (void)b_ih; (void)w_hh; (void)b_hh;  // consumed by gru_tc_prepare (packed into workspace)
  if (D != tc::kD) {
    set_error("tcgen05 engine: D must be 128, got %d", D);
    return DDFA_ERR_UNSUPPORTED;
  }
  if (workspace == nullptr || workspace_bytes < tc::kPackedBytes) {
    set_error("tcgen05 engine: workspace too small (%zu < %zu)", workspace_bytes, tc::kPackedBytes);
    return DDFA_ERR_WORKSPACE;
  }
  DDFA_CUDA(cudaFuncSetAttribute(tc::gru_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemAlloc));
  const int tiles = (N + tc::kTileM - 1) / tc::kTileM;
  tc::gru_tc_fwd_kernel<<<tiles, tc::kThreads, tc::kSmemAlloc, stream>>>(s, h, indptr, static_cast<const uint8_t *>(workspace), N, h_out,
                                                                        save_gates);
  DDFA_CHECK_LAUNCH("gru_tc_fwd_kernel");
  return DDFA_OK;
}

}  // namespace ddfa
