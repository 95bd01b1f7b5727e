// tcgen05 engine, forward GRU step v3 (D == 128) — weights in TENSOR MEMORY, activations streamed as the B operand.
//
// Replaces DGL GatedGraphConv's per-step `a = W h (summed over in-edges); h = GRUCell(a, h)` (reference:
// DDFA/code_gnn/models/flow_gnn/ggnn.py:60-63 -> dgl.nn.GatedGraphConv.forward -> torch.nn.GRUCell) for one step, given
// the edge-gathered sum s = A h and h as activation images:
//     gi = s W'^T + deg * b' + b_ih ,  gh = h Whh^T + b_hh ,  r,z = sigmoid(gi + gh) ,  n = tanh(gi_n + r * gh_n) ,
//     h' = n + z (h - n)
//
// Why v3.  The v2 kernel (gru_tc_fwd.cu) keeps a 96 KB weight slice in shared memory, which leaves two 32 KB operand
// stages; its timeline (profiles/r01l_trace_fwd.log) shows ~1 us per copy in flight, i.e. a feed of ~35 GB/s per SM, and
// copy_bench2 (profiles/r01m_copy_bench2.log) shows a B200 SM needs >= 3 x 64 KB in flight to pull > 100 GB/s.
// Here the GEMM is transposed,  D^T[gate column, node] = W[gate column, K] * X[node, K]^T :
//   * A = the weight slice, resident in TMEM for the life of the CTA: lane = one of the slice's 4 x 32 pre-activation
//     columns [gi_n | r | z | gh_n], K = 256 = [s part | h part] (zero blocks where a pre-activation does not use a part),
//     two bf16 per 32-bit column: 128 columns hi + 128 columns lo;
//   * B = the s and h image tiles (K-major SWIZZLE_128B, N = 128 nodes), THREE 64 KB stages, one bulk copy per tile;
//   * D = two 128-node accumulator buffers (TMEM columns 256..511): 48 MMAs (M = 128, N = 128, K = 16) per tile,
//     bf16x3: w_hi x_hi + w_lo x_hi + w_hi x_lo;
//   * epilogue: the four pre-activations of an output element sit in four TMEM lane quarters, i.e. four warps; they are
//     exchanged through a 16-node shared-memory tile, then a warp owns whole nodes: every global access is one full
//     128-byte line (32 consecutive columns of a node).
#include "tc_common.cuh"

namespace ddfa {
namespace tc3 {
using namespace tcc;

constexpr int kSlices = 4;
constexpr int kSliceCols = kD / kSlices;                  // 32 output columns per CTA
constexpr int kStages = 3;
constexpr int kStageBytes = kImageTileBytes;              // 64 KB: one operand tile [hi|lo][kb0|kb1]
constexpr int kWColsHalf = 128;                           // K = 256 bf16 -> 128 packed columns per variant
constexpr int kAccCol = 2 * kWColsHalf;                   // accumulators: TMEM columns 256 .. 511
constexpr int kChunks = 2 * kWColsHalf / 16;              // 16 chunks of 16 TMEM columns
constexpr int kXFloats = 4 * 16 * 32;                     // exchange tile: [pre-activation][16 nodes][32 columns]
constexpr int kOffX = kStages * kStageBytes;              // 192 KB
constexpr int kXBytes = 2 * 2 * kXFloats * 4;             // [node half][chunk parity] = 32 KB
constexpr int kOffBar = kOffX + kXBytes;
constexpr int kNumBars = 2 * kStages + 4 + 1;             // a_full, a_empty, acc_full[2], acc_empty[2], w_ready
constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr int kSmemAlloc = kOffTmemPtr + 16 + 1024;
constexpr int kThreads = 320;
constexpr int kEpiWarps = 8;
constexpr size_t kPackedWBytes = (size_t)kSlices * kChunks * 128 * 64;     // [slice][chunk][lane][16 words] = 512 KB
constexpr size_t kPackedBytes = kPackedWBytes + (size_t)kSlices * 128 * 8;  // + [slice][lane] {bias, degree bias}
static_assert(kSmemAlloc <= 232448, "shared memory budget");

// lane L of slice j: block = L / 32 (0 gi_n, 1 r, 2 z, 3 gh_n), output column oc = 32 j + L % 32.
// K index kk: part p = kk / 128 (0: s -> W', 1: h -> Whh), c = kk % 128.
__device__ __forceinline__ float weight_of(const float *__restrict__ w_fold, const float *__restrict__ w_hh, int blk, int oc, int kk) {
  const int p = kk >> 7, c = kk & 127;
  if (blk == 0) return p == 0 ? w_fold[(size_t)(2 * kD + oc) * kD + c] : 0.f;
  if (blk == 3) return p == 1 ? w_hh[(size_t)(2 * kD + oc) * kD + c] : 0.f;
  const int gate = blk - 1;
  return (p == 0 ? w_fold : w_hh)[(size_t)(gate * kD + oc) * kD + c];
}

__global__ void pack_kernel(const float *__restrict__ w_fold, const float *__restrict__ w_hh, const float *__restrict__ b_fold,
                            const float *__restrict__ b_ih, const float *__restrict__ b_hh, uint8_t *__restrict__ packed) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= kSlices * kChunks * 128) return;
  const int lane = idx % 128, chunk = (idx / 128) % kChunks, slice = idx / (128 * kChunks);
  const int blk = lane >> 5, oc = slice * kSliceCols + (lane & 31);
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int col = chunk * 16 + i;
    const int v = col >= kWColsHalf ? 1 : 0;
    const int kk = (col - v * kWColsHalf) * 2;
    __nv_bfloat16 h0, l0, h1, l1;
    split_bf16(weight_of(w_fold, w_hh, blk, oc, kk), h0, l0);
    split_bf16(weight_of(w_fold, w_hh, blk, oc, kk + 1), h1, l1);
    w[i] = v ? ((uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16))
             : ((uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16));
  }
  uint4 *dst = reinterpret_cast<uint4 *>(packed) + (size_t)idx * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  if (chunk == 0) {
    float2 b;
    if (blk == 0) b = make_float2(b_ih[2 * kD + oc], b_fold[2 * kD + oc]);
    else if (blk == 3) b = make_float2(b_hh[2 * kD + oc], 0.f);
    else b = make_float2(b_ih[(blk - 1) * kD + oc] + b_hh[(blk - 1) * kD + oc], b_fold[(blk - 1) * kD + oc]);
    reinterpret_cast<float2 *>(packed + kPackedWBytes)[slice * 128 + lane] = b;
  }
}

__global__ void __launch_bounds__(kThreads, 1) gru_fwd3_kernel(const uint8_t *__restrict__ s_img, const uint8_t *__restrict__ h_img,
                                                               const float *__restrict__ h, const int32_t *__restrict__ indptr,
                                                               const uint8_t *__restrict__ packed, int32_t N,
                                                               float *__restrict__ h_out, uint8_t *__restrict__ h_out_img,
                                                               float *__restrict__ gates) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + kOffBar;
  auto a_full = [&](int i) { return bar0 + 8u * i; };
  auto a_empty = [&](int i) { return bar0 + 8u * (kStages + i); };
  auto acc_full = [&](int i) { return bar0 + 8u * (2 * kStages + i); };
  auto acc_empty = [&](int i) { return bar0 + 8u * (2 * kStages + 2 + i); };
  const uint32_t w_ready = bar0 + 8u * (2 * kStages + 4);
  volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(smem + kOffTmemPtr);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x % kSlices;
  const int group = blockIdx.x / kSlices, num_groups = gridDim.x / kSlices;
  const int num_tiles = (N + kTileM - 1) / kTileM;
  const int my_tiles = (num_tiles > group) ? (num_tiles - 1 - group) / num_groups + 1 : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(a_full(i), 1); mbar_init(a_empty(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full(i), 1); mbar_init(acc_empty(i), kEpiWarps); }
    mbar_init(w_ready, kEpiWarps);
    mbar_fence_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(smem_u32((const void *)tmem_ptr_smem), 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int tron = g_trace_on;
  if (threadIdx.x == 0) trace_stamp(tron, 0, 0);

  if (warp == 0) {
    // ===== producer: per tile the s tile and the h tile, one 64 KB bulk copy each =====
    if (my_tiles > 0 && elect_one()) {
      int cc = 0;
      for (int k = 0; k < my_tiles; ++k) {
        const int tile = group + k * num_groups;
        for (int p = 0; p < 2; ++p, ++cc) {
          const int stage = cc % kStages, use = cc / kStages;
          if (use > 0) mbar_wait(a_empty(stage), (use - 1) & 1);
          if (p == 0) trace_stamp(tron, k, 1);
          mbar_arrive_expect_tx(a_full(stage), kStageBytes);
          bulk_g2s(sbase + stage * kStageBytes, (p == 0 ? s_img : h_img) + (size_t)tile * kImageTileBytes, kStageBytes, a_full(stage));
          if (p == 1) trace_stamp(tron, k, 2);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (my_tiles > 0 && elect_one()) {
      constexpr uint32_t kIdesc128 = make_idesc(128);
      mbar_wait(w_ready, 0);
      tc_fence_after();
      int cc = 0;
      for (int k = 0; k < my_tiles; ++k) {
        const int buf = k & 1, buse = k >> 1;
        if (buse > 0) mbar_wait(acc_empty(buf), (buse - 1) & 1);
        tc_fence_after();
        trace_stamp(tron, k, 3);
        const uint32_t d_addr = tmem_base + (uint32_t)(kAccCol + buf * 128);
        for (int p = 0; p < 2; ++p, ++cc) {
          const int stage = cc % kStages, use = cc / kStages;
          mbar_wait(a_full(stage), use & 1);
          tc_fence_after();
          if (p == 0) trace_stamp(tron, k, 4);
          if (p == 1) trace_stamp(tron, k, 5);
          const uint64_t b_base = make_desc(sbase + stage * kStageBytes);     // chunks [hi kb0 | hi kb1 | lo kb0 | lo kb1]
          const uint32_t a_base = tmem_base + (uint32_t)(p * 64);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint32_t kk2 = (uint32_t)(kb * 32 + k4 * 8);              // packed weight column of this K step
              const uint64_t b_hi = desc_advance(b_base, (uint32_t)kb * kChunkBytes + k4 * 32);
              const uint64_t b_lo = desc_advance(b_base, (uint32_t)(2 + kb) * kChunkBytes + k4 * 32);
              const uint32_t first = (p == 0 && kb == 0 && k4 == 0) ? 0u : 1u;
              umma_f16_ts(d_addr, a_base + kk2, b_hi, kIdesc128, first);                    // w_hi x_hi
              umma_f16_ts(d_addr, a_base + kWColsHalf + kk2, b_hi, kIdesc128, 1u);          // w_lo x_hi
              umma_f16_ts(d_addr, a_base + kk2, b_lo, kIdesc128, 1u);                       // w_hi x_lo
            }
          }
          umma_commit(a_empty(stage));
        }
        umma_commit(acc_full(buf));
        trace_stamp(tron, k, 6);
      }
    }
  } else {
    // ===== weights -> tensor memory, then the epilogue =====
    const int q = warp & 3;                  // TMEM lane quarter = pre-activation: 0 gi_n, 1 r, 2 z, 3 gh_n
    const int hh = (warp - 2) >> 2;          // node half of the tile: nodes 64 hh .. 64 hh + 63
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int gcol = slice * kSliceCols + lane;       // this lane's output column of h
    float2 bias = make_float2(0.f, 0.f);
    if (my_tiles > 0) {
      const uint4 *src = reinterpret_cast<const uint4 *>(packed) + ((size_t)slice * kChunks * 128 + (size_t)(q * 32 + lane)) * 4;
#pragma unroll 4
      for (int c = 0; c < kChunks / 2; ++c) {
        const int chunk = hh * (kChunks / 2) + c;
        const uint4 *p = src + (size_t)chunk * 128 * 4;
        const uint4 x0 = __ldg(p), x1 = __ldg(p + 1), x2 = __ldg(p + 2), x3 = __ldg(p + 3);
        const uint32_t w[16] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w, x2.x, x2.y, x2.z, x2.w, x3.x, x3.y, x3.z, x3.w};
        tmem_st16(lane_addr + (uint32_t)(chunk * 16), w);
      }
      tmem_st_wait();
      bias = __ldg(reinterpret_cast<const float2 *>(packed + kPackedWBytes) + slice * 128 + q * 32 + lane);
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(w_ready);

    const size_t plane = (size_t)N * kD;
    float *X = reinterpret_cast<float *>(smem + kOffX) + (size_t)hh * 2 * kXFloats;
    const int bar_id = 1 + hh;
    const bool tr = (warp == 2 && lane == 0);
    const bool is_sigmoid = (q == 1 || q == 2);
    for (int k = 0; k < my_tiles; ++k) {
      const int tile = group + k * num_groups;
      const int buf = k & 1, buse = k >> 1;
      const int64_t node_h = (int64_t)tile * kTileM + hh * 64;          // first node of this warp's half
      if (tr) trace_stamp(tron, k, 7);
      // h of the 16 nodes this warp finishes (4 per 16-node chunk), fetched while the MMAs of the tile run
      float hp[16];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t node = node_h + c * 16 + q * 4 + j;
          hp[c * 4 + j] = node < N ? __ldg(h + node * kD + gcol) : 0.f;
        }
      // in-degrees of the half's 64 nodes: lane l holds deg(node_h + l) and deg(node_h + 32 + l)
      float dlo, dhi;
      {
        auto ip = [&](int64_t i) { return __ldg(indptr + (i < N ? i : (int64_t)N)); };
        const int a0 = ip(node_h + lane), a1 = ip(node_h + 32 + lane), a2 = ip(node_h + 64);
        const int n0 = __shfl_down_sync(0xffffffffu, a0, 1), n1 = __shfl_down_sync(0xffffffffu, a1, 1);
        const int a1_0 = __shfl_sync(0xffffffffu, a1, 0);
        dlo = (float)((lane < 31 ? n0 : a1_0) - a0);
        dhi = (float)((lane < 31 ? n1 : a2) - a1);
      }
      mbar_wait(acc_full(buf), buse & 1);
      tc_fence_after();
      if (tr) trace_stamp(tron, k, 8);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v[16];
        tmem_ld16(lane_addr + (uint32_t)(kAccCol + buf * 128 + hh * 64 + c * 16), v);
        tmem_ld_wait();
        if (c == 3) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty(buf));   // this warp has read its part of the accumulator buffer
          if (tr) trace_stamp(tron, k, 9);
        }
        float *Xc = X + (c & 1) * kXFloats;
        const float dsel = (c < 2) ? dlo : dhi;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float deg = __shfl_sync(0xffffffffu, dsel, (c & 1) * 16 + i);
          const float pre = v[i] + fmaf(deg, bias.y, bias.x);
          Xc[(q * 16 + i) * 32 + lane] = is_sigmoid ? fast_sigmoid(pre) : pre;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        // this warp now owns nodes 4q .. 4q+3 of the chunk, lane = column
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = q * 4 + j;
          const int64_t node = node_h + c * 16 + i;
          const bool valid = node < N;
          const float gin = Xc[(0 * 16 + i) * 32 + lane], r = Xc[(1 * 16 + i) * 32 + lane];
          const float z = Xc[(2 * 16 + i) * 32 + lane], ghn = Xc[(3 * 16 + i) * 32 + lane];
          const float n = fast_tanh(fmaf(r, ghn, gin));
          const float hnew = valid ? fmaf(z, hp[c * 4 + j] - n, n) : 0.f;   // rows past N stay zero in the image
          if (valid) {
            h_out[node * kD + gcol] = hnew;
            if (gates) {
              float *g0 = gates + node * kD + gcol;
              g0[0] = r;
              g0[plane] = z;
              g0[2 * plane] = n;
              g0[3 * plane] = ghn;
            }
          }
          if (h_out_img) {
            // lanes 0-15 write the hi words, 16-31 the lo words of column pairs (2 pi, 2 pi + 1)
            const int pi = lane & 15;
            const float x0 = __shfl_sync(0xffffffffu, hnew, 2 * pi), x1 = __shfl_sync(0xffffffffu, hnew, 2 * pi + 1);
            __nv_bfloat16 h0, l0, h1, l1;
            split_bf16(x0, h0, l0);
            split_bf16(x1, h1, l1);
            const uint32_t word = (lane < 16) ? ((uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16))
                                              : ((uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16));
            *reinterpret_cast<uint32_t *>(h_out_img + image_offset(node, slice * kSliceCols + 2 * pi, lane >> 4)) = word;
          }
        }
      }
      if (tr) trace_stamp(tron, k, 10);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace tc3

size_t gru_tc3_packed_bytes() { return tc3::kPackedBytes; }

int gru_tc3_prepare(const float *w_fold, const float *b_fold, const float *b_ih, const float *w_hh, const float *b_hh, void *packed,
                    cudaStream_t stream) {
  const int total = tc3::kSlices * tc3::kChunks * 128;
  tc3::pack_kernel<<<(total + 127) / 128, 128, 0, stream>>>(w_fold, w_hh, b_fold, b_ih, b_hh, static_cast<uint8_t *>(packed));
  DDFA_CHECK_LAUNCH("tc3::pack_kernel");
  return DDFA_OK;
}

int gru_tc3_step_fwd(const void *s_img, const void *h_img, const float *h, const int32_t *indptr, int32_t N, float *h_out,
                     void *h_out_img, float *save_gates, const void *packed, cudaStream_t stream) {
  const int tiles = (N + tcc::kTileM - 1) / tcc::kTileM;
  DDFA_CUDA(cudaFuncSetAttribute(tc3::gru_fwd3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc3::kSmemAlloc));
  int groups = kNumSMs / tc3::kSlices;
  if (groups > tiles) groups = tiles;
  tc3::gru_fwd3_kernel<<<groups * tc3::kSlices, tc3::kThreads, tc3::kSmemAlloc, stream>>>(
      static_cast<const uint8_t *>(s_img), static_cast<const uint8_t *>(h_img), h, indptr, static_cast<const uint8_t *>(packed), N, h_out,
      static_cast<uint8_t *>(h_out_img), save_gates);
  DDFA_CHECK_LAUNCH("tc3::gru_fwd3_kernel");
  return DDFA_OK;
}

int gru_tc3_trace_enable(int on) {
  DDFA_CUDA(cudaMemcpyToSymbol(tcc::g_trace_on, &on, sizeof(int)));
  return DDFA_OK;
}
int gru_tc3_trace_read(void *host, size_t bytes) {
  if (bytes > tcc::kTraceWords * sizeof(long long)) bytes = tcc::kTraceWords * sizeof(long long);
  DDFA_CUDA(cudaMemcpyFromSymbol(host, tcc::g_trace, bytes));
  return DDFA_OK;
}

}  // namespace ddfa
