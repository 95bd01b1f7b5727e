// tcgen05 engine, forward GRU step v2 (D == 128) — weight-stationary, persistent, TMA-fed.
//
//   acc_r = s W'_r^T + h Whh_r^T    acc_z = s W'_z^T + h Whh_z^T    acc_gin = s W'_n^T    acc_ghn = h Whh_n^T
//   r,z = sigmoid(acc + indeg b' + b_ih + b_hh) ; n = tanh(gin + r * ghn) ; h' = n + z (h - n)
//
// Why this shape (r01c ncu: the first tcgen05 kernel kept the tensor pipe busy only 7-10 % of the time —
// every CTA re-streamed all 384 KB of split weights per 128-node tile and converted its A operand itself,
// with load / MMA / epilogue phases serialised):
//   * WEIGHT-STATIONARY: a CTA owns a 32-column slice of the 128 hidden columns for the whole launch; the
//     bf16 hi/lo images of the 96 weight rows it needs (3 gates x 32 columns, K = 128, two matrices) are
//     96 KB and are loaded into shared memory ONCE per CTA.
//   * The A operands (s and h) arrive as "activation images" (tc_common.cuh) written by the producer
//     kernels, so the GEMM kernel has no conversion pass: one elected thread streams 16 KB chunks through a
//     7-stage ring with cp.async.bulk (TMA 1-D bulk copy, SASS UBLKCP) + mbarrier complete_tx.
//   * The accumulators of one 128-node tile are only 4 x 32 = 128 TMEM columns, so FOUR tiles are in flight
//     in the 512-column TMEM: the epilogue of tile k overlaps the MMAs of tiles k+1..k+3.
//   * PERSISTENT: grid = 4 slices x up to 37 tile groups (148 SMs); group g walks tiles g, g+G, g+2G, ...
//     The four slice-CTAs of a group read the same A chunks at the same time (L2 hits).
// Precision: bf16x3 (a_hi w_hi + a_hi w_lo + a_lo w_hi, fp32 accumulate in TMEM), gate math with ex2.approx.
// Warp roles: warp 0 = TMEM alloc + TMA producer, warp 1 = MMA issuer (one lane), warps 2..9 = epilogue
// (thread = node row, 16 of the slice's 32 columns each).
#include "tc_common.cuh"

namespace ddfa {
namespace tc2 {
using namespace tcc;

constexpr int kSlices = 4;
constexpr int kSliceCols = kD / kSlices;                 // 32
constexpr int kWRows = 3 * kSliceCols;                   // 96 weight rows per slice: [r | z | n]
constexpr int kWImgBytes = kWRows * 128;                 // 12 KB: one (matrix p, kblock, variant) image
constexpr int kWSliceBytes = 8 * kWImgBytes;             // 96 KB per slice, index (p*2 + kb)*2 + v
constexpr int kBiasSlice = 7 * kSliceCols;               // floats per slice
// A feed: cp.async.bulk of one VARIANT (hi or lo) of an operand tile = 32 KB ([kb0 | kb1], contiguous in the image),
// two stages.  profiles/r01g_copy_bench.log: a 1-D bulk copy costs ~0.42 us almost independently of its size and the
// copies of one SM do not overlap, so throughput = size / 0.42 us (16 KB -> 38 GB/s/SM, 32 KB -> 70 GB/s/SM).
constexpr int kAStages = 2;
constexpr int kAStageBytes = 2 * kChunkBytes;            // 32 KB
constexpr int kAccBufs = 4;
constexpr int kOffA = kWSliceBytes;
// epilogue staging (tc_common.cuh "coalesced epilogue I/O"): per warp pair 2 output planes + 1 input plane
constexpr int kStgPlanes = 3;
constexpr int kOffStage = kOffA + kAStages * kAStageBytes;
constexpr int kStageBytes = 4 * kStgPlanes * kStagePlaneFloats * 4;
constexpr int kOffBias = kOffStage + kStageBytes;
constexpr int kOffBar = kOffBias + kBiasSlice * 4;
constexpr int kNumBars = 1 + 3 * kAStages + 2 * kAccBufs;   // w_full, a_full/a_empty/a_ready[stages], acc_full/acc_empty[bufs]
constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr int kSmemAlloc = kOffTmemPtr + 16 + 1024;
constexpr int kThreads = 320;
constexpr int kEpiWarps = 8;
constexpr size_t kPackedBytes = (size_t)kSlices * kWSliceBytes + (size_t)kSlices * kBiasSlice * 4;

// ---- fp32 [N,128] -> activation image (zero tail rows) ----------------------------------------
__global__ void __launch_bounds__(256) to_image_kernel(const float *__restrict__ x, int32_t N, uint8_t *__restrict__ image) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // one thread = 8 consecutive columns of a row
  const int64_t rows = ((int64_t)N + kTileM - 1) / kTileM * kTileM;
  if (t >= rows * 16) return;
  const int64_t node = t >> 4;
  const int col = (int)(t & 15) * 8;
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (node < N) {
    const float4 a = ldg_nc_f4(x + node * kD + col), b = ldg_nc_f4(x + node * kD + col + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  uint4 ph, pl;
  split8(v, ph, pl);
  *reinterpret_cast<uint4 *>(image + image_offset(node, col, 0)) = ph;
  *reinterpret_cast<uint4 *>(image + image_offset(node, col, 1)) = pl;
}

// ---- per-slice weight images + biases ----------------------------------------------------------
// packed = [slice j][ (p*2+kb)*2+v ][96 rows x 64 k swizzled]  then  [slice j][7][32] biases
// local row lr = gate*32 + c  <->  weight row gate*128 + 32 j + c ; p: 0 = w_fold (s part), 1 = w_hh (h part)
__global__ void __launch_bounds__(256) pack_kernel(const float *__restrict__ w_fold, const float *__restrict__ w_hh,
                                                   const float *__restrict__ b_fold, const float *__restrict__ b_ih,
                                                   const float *__restrict__ b_hh, uint8_t *__restrict__ packed) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = kSlices * 2 * 2 * kWRows * 8;  // (j, p, kb, lr, k8)
  if (t < total) {
    const int k8 = t & 7;
    const int lr = (t >> 3) % kWRows;
    const int rest = (t >> 3) / kWRows;  // (j*2 + p)*2 + kb
    const int kb = rest & 1, p = (rest >> 1) & 1, j = rest >> 2;
    // row order inside an image: s part (p = 0) [n | r | z], h part (p = 1) [r | z | n] — so that each part is ONE
    // N = 96 MMA into the TMEM column layout [gi_n | r | z | gh_n] (p = 0 -> columns 0..95, p = 1 -> columns 32..127)
    const int blk = lr / kSliceCols, c = lr % kSliceCols;
    const int gate = (p == 0) ? (blk == 0 ? 2 : blk - 1) : blk;
    const float *W = (p == 0 ? w_fold : w_hh) + (size_t)(gate * kD + j * kSliceCols + c) * kD + kb * 64 + k8 * 8;
    const float4 a = *reinterpret_cast<const float4 *>(W), b = *reinterpret_cast<const float4 *>(W + 4);
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint4 ph, pl;
    split8(x, ph, pl);
    uint8_t *base = packed + (size_t)j * kWSliceBytes + (size_t)((p * 2 + kb) * 2) * kWImgBytes + sw128_offset(lr, k8 * 8);
    *reinterpret_cast<uint4 *>(base) = ph;
    *reinterpret_cast<uint4 *>(base + kWImgBytes) = pl;
  }
  if (t < kD) {
    const int j = t / kSliceCols, c = t % kSliceCols;
    float *bias = reinterpret_cast<float *>(packed + (size_t)kSlices * kWSliceBytes) + j * kBiasSlice;
    bias[0 * kSliceCols + c] = b_ih[t] + b_hh[t];
    bias[1 * kSliceCols + c] = b_ih[kD + t] + b_hh[kD + t];
    bias[2 * kSliceCols + c] = b_ih[2 * kD + t];
    bias[3 * kSliceCols + c] = b_hh[2 * kD + t];
    bias[4 * kSliceCols + c] = b_fold[t];
    bias[5 * kSliceCols + c] = b_fold[kD + t];
    bias[6 * kSliceCols + c] = b_fold[2 * kD + t];
  }
}

// CLUSTER: the 4 slice-CTAs of a tile group form a thread-block cluster and each operand variant of a tile (32 KB) is
// fetched from L2 ONCE — CTA pv issues variant pv (s_hi, s_lo, h_hi, h_lo) as a multicast bulk copy into the same stage of
// all four CTAs, after every CTA has armed its full barrier for that stage use and signalled it on the issuer's a_ready
// barrier.  Per-SM copy-engine work drops from 4 copies per tile to 1 (the engine serialises copies at ~0.45 us each).
template <bool CLUSTER>
__global__ void __launch_bounds__(kThreads, 1) gru_fwd_kernel(const uint8_t *__restrict__ s_img, const uint8_t *__restrict__ h_img,
                                                              const float *__restrict__ h, const int32_t *__restrict__ indptr,
                                                              const uint8_t *__restrict__ packed, int32_t N,
                                                              float *__restrict__ h_out, uint8_t *__restrict__ h_out_img,
                                                              float *__restrict__ gates) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + kOffBar;
  const uint32_t w_full = bar0;
  auto a_full = [&](int i) { return bar0 + 8u * (1 + i); };
  auto a_empty = [&](int i) { return bar0 + 8u * (1 + kAStages + i); };
  auto acc_full = [&](int i) { return bar0 + 8u * (1 + 2 * kAStages + i); };
  auto acc_empty = [&](int i) { return bar0 + 8u * (1 + 2 * kAStages + kAccBufs + i); };
  auto a_ready = [&](int i) { return bar0 + 8u * (1 + 2 * kAStages + 2 * kAccBufs + i); };   // CLUSTER only
  volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(smem + kOffTmemPtr);
  float *bias_s = reinterpret_cast<float *>(smem + kOffBias);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x % kSlices;
  const int group = blockIdx.x / kSlices, num_groups = gridDim.x / kSlices;
  const int num_tiles = (N + kTileM - 1) / kTileM;
  const int my_tiles = (num_tiles > group) ? (num_tiles - 1 - group) / num_groups + 1 : 0;

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < kAStages; ++i) { mbar_init(a_full(i), 1); mbar_init(a_empty(i), 1); }
    for (int i = 0; i < kAccBufs; ++i) { mbar_init(acc_full(i), 1); mbar_init(acc_empty(i), kEpiWarps); }
    for (int i = 0; i < kAStages; ++i) mbar_init(a_ready(i), kSlices);
    mbar_fence_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(smem_u32((const void *)tmem_ptr_smem), 512);
  }
  {
    const float *bias_g = reinterpret_cast<const float *>(packed + (size_t)kSlices * kWSliceBytes) + slice * kBiasSlice;
    for (int i = threadIdx.x; i < kBiasSlice; i += kThreads) bias_s[i] = bias_g[i];
  }
  tc_fence_before();
  __syncthreads();
  if (CLUSTER) cluster_sync_all();   // every CTA's mbarriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int tron = g_trace_on;
  if (threadIdx.x == 0) trace_stamp(tron, 0, 0);

  if (warp == 0) {
    // ===== TMA producer: resident weights once, then per tile the operand variants s_hi, s_lo, h_hi, h_lo =====
    if (my_tiles > 0 && elect_one()) {
      mbar_arrive_expect_tx(w_full, kWSliceBytes);
      bulk_g2s(sbase, packed + (size_t)slice * kWSliceBytes, kWSliceBytes, w_full);   // one 96 KB copy (copy cost ~ size-independent)
      int cc = 0;
      for (int k = 0; k < my_tiles; ++k) {
        const int tile = group + k * num_groups;
        for (int pv = 0; pv < 4; ++pv, ++cc) {
          const int p = pv >> 1, v = pv & 1;
          const int stage = cc % kAStages, use = cc / kAStages;
          if (use > 0) mbar_wait(a_empty(stage), (use - 1) & 1);
          if (pv == 0) trace_stamp(tron, k, 1);
          mbar_arrive_expect_tx(a_full(stage), kAStageBytes);
          const uint8_t *src = (p == 0 ? s_img : h_img) + (size_t)tile * kImageTileBytes + (size_t)v * kAStageBytes;
          if (!CLUSTER) {
            bulk_g2s(sbase + kOffA + stage * kAStageBytes, src, kAStageBytes, a_full(stage));
          } else {
            // tell the issuer of this variant (CTA rank pv) that my stage is free and armed; the issuer waits for all 4
            mbar_arrive_remote(a_ready(stage), (uint32_t)pv);
            if (slice == pv) {
              mbar_wait_cluster(a_ready(stage), k & 1);     // this CTA issues once per tile on stage pv % 2
              bulk_g2s_mcast(sbase + kOffA + stage * kAStageBytes, src, kAStageBytes, a_full(stage), (uint16_t)0xF);
            }
          }
          if (pv == 3) trace_stamp(tron, k, 2);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (my_tiles > 0 && elect_one()) {
      constexpr uint32_t kIdesc96 = make_idesc(96);
      mbar_wait(w_full, 0);
      int cc = 0;
      for (int k = 0; k < my_tiles; ++k) {
        const int buf = k % kAccBufs, buse = k / kAccBufs;
        if (buse > 0) mbar_wait(acc_empty(buf), (buse - 1) & 1);
        tc_fence_after();
        trace_stamp(tron, k, 3);
        const uint32_t d_base = tmem_base + (uint32_t)buf * 128u;   // [gin 0-31 | r 32-63 | z 64-95 | ghn 96-127]
        for (int pv = 0; pv < 4; ++pv, ++cc) {
          const int p = pv >> 1, v = pv & 1;
          const int stage = cc % kAStages, use = cc / kAStages;
          mbar_wait(a_full(stage), use & 1);
          tc_fence_after();
          if (pv == 0) trace_stamp(tron, k, 4);
          if (pv == 3) trace_stamp(tron, k, 5);
          const int n_wv = (v == 0) ? 2 : 1;   // a_hi pairs with w_hi and w_lo; a_lo with w_hi only
          for (int kb = 0; kb < 2; ++kb) {
            const uint32_t a_addr = sbase + kOffA + stage * kAStageBytes + (uint32_t)kb * kChunkBytes;
            for (int wv = 0; wv < n_wv; ++wv) {
              const uint32_t w_addr = sbase + (uint32_t)(((p * 2 + kb) * 2 + wv) * kWImgBytes);
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                const uint64_t ad = make_desc(a_addr + k4 * 32);
                const bool first = (v == 0 && kb == 0 && wv == 0 && k4 == 0);
                // p = 0: [gin | r | z] (zero-initialised by the first MMA of the tile);
                // p = 1: [r | z | ghn] accumulates onto r, z — ghn must start from zero, so the first p = 1 MMA of a tile
                //        is split into an accumulating N = 64 and a zero-initialising N = 32.
                if (p == 0) {
                  umma_f16(d_base, ad, make_desc(w_addr + k4 * 32), kIdesc96, first ? 0u : 1u);
                } else if (!first) {
                  umma_f16(d_base + 32, ad, make_desc(w_addr + k4 * 32), kIdesc96, 1u);
                } else {
                  umma_f16(d_base + 32, ad, make_desc(w_addr + k4 * 32), make_idesc(64), 1u);
                  umma_f16(d_base + 96, ad, make_desc(w_addr + 64 * 128 + k4 * 32), make_idesc(32), 0u);
                }
              }
            }
          }
          umma_commit(a_empty(stage));
        }
        umma_commit(acc_full(buf));
        trace_stamp(tron, k, 6);
      }
    }
  } else {
    // ===== epilogue: warps w and w+4 share a TMEM lane quarter (32 rows) and hold column halves 0 / 1 of the slice =====
    const int lw = warp - 2;
    const int q = warp & 3;          // TMEM lane quarter
    const int csub = lw >> 2;        // which 16 of the slice's 32 columns
    const int bar_id = 1 + q;        // named barrier of this warp pair
    float *P0 = reinterpret_cast<float *>(smem + kOffStage) + (size_t)q * kStgPlanes * kStagePlaneFloats;
    float *P1 = P0 + kStagePlaneFloats, *PI = P1 + kStagePlaneFloats;
    const int lc0 = csub * 16;                         // first column inside the slice
    const int gcs = slice * kSliceCols;                // first global column of the slice
    const size_t plane = (size_t)N * kD;
    auto rows_of = [&](int kk) -> int {                // valid rows of this pair's 32-row block in tile kk
      if (kk >= my_tiles) return 0;
      const int64_t r0 = (int64_t)(group + kk * num_groups) * kTileM + q * 32;
      const int64_t rem = (int64_t)N - r0;
      return rem <= 0 ? 0 : (rem > 32 ? 32 : (int)rem);
    };
    float4 hreg[4];                                    // next tile's h rows, fetched one tile ahead (coalesced)
    {
      const int64_t r0 = (int64_t)group * kTileM + q * 32;
      stage_fetch_rows(h + r0 * kD + gcs, kD, lane, csub, rows_of(0), hreg);
    }
    for (int k = 0; k < my_tiles; ++k) {
      const int tile = group + k * num_groups;
      const int buf = k % kAccBufs, buse = k / kAccBufs;
      const int64_t r0 = (int64_t)tile * kTileM + q * 32;      // first node row of this pair's block
      const int64_t node = r0 + lane;
      const int rows_valid = rows_of(k);
      const bool valid = lane < rows_valid;
      const bool tr = (warp == 2 && lane == 0);
      if (tr) trace_stamp(tron, k, 7);
      // h rows of this tile: registers -> staging -> this thread's row piece
      stage_put_rows(PI, lane, csub, hreg);
      pair_sync(bar_id);
      float hv[16];
      stage_read16(PI, lane, csub, hv);
      {
        const int64_t rn = (int64_t)(group + (k + 1) * num_groups) * kTileM + q * 32;
        stage_fetch_rows(h + rn * kD + gcs, kD, lane, csub, rows_of(k + 1), hreg);
      }
      const float deg = valid ? (float)(__ldg(indptr + node + 1) - __ldg(indptr + node)) : 0.f;

      mbar_wait(acc_full(buf), buse & 1);
      tc_fence_after();
      if (tr) trace_stamp(tron, k, 8);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 128 + lc0);
      float va[16], vb[16];
      tmem_ld16(taddr + 32, va);   // r accumulator
      tmem_ld16(taddr + 96, vb);   // gh_n accumulator
      tmem_ld_wait();
      float o_r[16], o_g[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = lc0 + i;
        o_r[i] = fast_sigmoid(va[i] + fmaf(deg, bias_s[4 * kSliceCols + c], bias_s[0 * kSliceCols + c]));
        o_g[i] = vb[i] + bias_s[3 * kSliceCols + c];
      }
      tmem_ld16(taddr + 64, va);   // z accumulator
      tmem_ld16(taddr + 0, vb);    // gi_n accumulator
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));   // this warp has drained its part of the buffer
      if (tr) trace_stamp(tron, k, 9);
      float o_z[16], o_n[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = lc0 + i;
        o_z[i] = fast_sigmoid(va[i] + fmaf(deg, bias_s[5 * kSliceCols + c], bias_s[1 * kSliceCols + c]));
        o_n[i] = fast_tanh(vb[i] + fmaf(deg, bias_s[6 * kSliceCols + c], bias_s[2 * kSliceCols + c]) + o_r[i] * o_g[i]);
      }
      if (gates) {
        float *g0 = gates + r0 * kD + gcs;
        stage_write16(P0, lane, csub, o_r);
        stage_write16(P1, lane, csub, o_g);
        pair_sync(bar_id);
        stage_store_rows(P0, g0, kD, lane, csub, rows_valid);
        stage_store_rows(P1, g0 + 3 * plane, kD, lane, csub, rows_valid);
        pair_sync(bar_id);
        stage_write16(P0, lane, csub, o_z);
        stage_write16(P1, lane, csub, o_n);
        pair_sync(bar_id);
        stage_store_rows(P0, g0 + plane, kD, lane, csub, rows_valid);
        stage_store_rows(P1, g0 + 2 * plane, kD, lane, csub, rows_valid);
        pair_sync(bar_id);
      }
      // h' (rows past N are zero: the image's padding rows must be zero for the weight-gradient GEMM)
      float o_h[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) o_h[i] = valid ? fmaf(o_z[i], hv[i] - o_n[i], o_n[i]) : 0.f;
      stage_write16(P0, lane, csub, o_h);
      if (h_out_img) {
        // image words of this thread's 16 columns: P1 row = [16 words hi | 16 words lo] for the slice's 32 columns
        uint4 ph0, pl0, ph1, pl1;
        {
          float x8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) x8[i] = o_h[i];
          split8(x8, ph0, pl0);
#pragma unroll
          for (int i = 0; i < 8; ++i) x8[i] = o_h[8 + i];
          split8(x8, ph1, pl1);
        }
        uint4 *rowp = reinterpret_cast<uint4 *>(P1 + lane * kStageLd);
        rowp[2 * csub] = ph0; rowp[2 * csub + 1] = ph1;
        rowp[4 + 2 * csub] = pl0; rowp[4 + 2 * csub + 1] = pl1;
      }
      pair_sync(bar_id);
      stage_store_rows(P0, h_out + r0 * kD + gcs, kD, lane, csub, rows_valid);
      if (h_out_img) {
        const int piece = lane & 7;              // 0-3: hi pieces (8 columns each), 4-7: lo pieces
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int row = csub * 16 + j * 4 + (lane >> 3);
          const uint4 w = *reinterpret_cast<const uint4 *>(P1 + row * kStageLd + piece * 4);
          *reinterpret_cast<uint4 *>(h_out_img + image_offset(r0 + row, gcs + 8 * (piece & 3), piece >> 2)) = w;
        }
      }
      pair_sync(bar_id);
      if (tr) trace_stamp(tron, k, 10);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CLUSTER) cluster_sync_all();   // no CTA may exit while peers can still multicast into it / arrive on its barriers
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace tc2

size_t act_image_bytes(int64_t n) { return tcc::image_bytes(n); }

int act_to_image(const float *x, int32_t N, void *image, cudaStream_t stream) {
  const int64_t rows = ((int64_t)N + tcc::kTileM - 1) / tcc::kTileM * tcc::kTileM;
  const int64_t total = rows * 16;
  if (total == 0) return DDFA_OK;
  tc2::to_image_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(x, N, static_cast<uint8_t *>(image));
  DDFA_CHECK_LAUNCH("to_image_kernel");
  return DDFA_OK;
}

// workspace = [v2 weight slices + biases][v3 packed weights + biases (gru_tc_fwd3.cu)]
static int g_fwd_v3 = 1;   // ddfa_debug_set key 4: 1 = weights-in-TMEM gru_fwd3_kernel, 0 = weight-slices-in-smem gru_fwd_kernel
void gru_tc2_set_fwd3(int on) { g_fwd_v3 = on; }
size_t gru_tc2_workspace_bytes() { return tc2::kPackedBytes + gru_tc3_packed_bytes(); }

int gru_tc2_prepare(const float *w_fold, const float *b_fold, const float *b_ih, const float *w_hh, const float *b_hh,
                    void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (workspace == nullptr || workspace_bytes < gru_tc2_workspace_bytes()) {
    set_error("tcgen05 engine: workspace too small (%zu < %zu)", workspace_bytes, gru_tc2_workspace_bytes());
    return DDFA_ERR_WORKSPACE;
  }
  if (g_fwd_v3) return gru_tc3_prepare(w_fold, b_fold, b_ih, w_hh, b_hh, static_cast<uint8_t *>(workspace) + tc2::kPackedBytes, stream);
  const int total = tc2::kSlices * 2 * 2 * tc2::kWRows * 8;
  tc2::pack_kernel<<<(total + 255) / 256, 256, 0, stream>>>(w_fold, w_hh, b_fold, b_ih, b_hh, static_cast<uint8_t *>(workspace));
  DDFA_CHECK_LAUNCH("tc2::pack_kernel");
  return DDFA_OK;
}

// tuning knob (ddfa_debug_set key 1): 1 = cluster multicast feed, 0 = every CTA copies for itself.  Measured on B200 at
// C0 (profiles/r01k_tcdebug.log): multicast 56.8 us vs unicast 43.3 us per launch, bit-identical results — at cluster
// size 4 the L2 already de-duplicates the four unicast reads, and the cluster handshake only adds lock-step latency.
static int g_fwd_cluster = 0;
void gru_tc2_set_cluster(int on) { g_fwd_cluster = on; }

int gru_tc2_trace_enable(int on) {
  DDFA_CUDA(cudaMemcpyToSymbol(tcc::g_trace_on, &on, sizeof(int)));
  return DDFA_OK;
}
int gru_tc2_trace_read(void *host, size_t bytes) {
  if (bytes > tcc::kTraceWords * sizeof(long long)) bytes = tcc::kTraceWords * sizeof(long long);
  DDFA_CUDA(cudaMemcpyFromSymbol(host, tcc::g_trace, bytes));
  return DDFA_OK;
}

int gru_tc2_step_fwd(const void *s_img, const void *h_img, const float *h, const int32_t *indptr, int32_t N, float *h_out,
                     void *h_out_img, float *save_gates, const void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (workspace == nullptr || workspace_bytes < gru_tc2_workspace_bytes()) {
    set_error("tcgen05 engine: workspace too small (%zu < %zu)", workspace_bytes, gru_tc2_workspace_bytes());
    return DDFA_ERR_WORKSPACE;
  }
  if (g_fwd_v3)
    return gru_tc3_step_fwd(s_img, h_img, h, indptr, N, h_out, h_out_img, save_gates,
                            static_cast<const uint8_t *>(workspace) + tc2::kPackedBytes, stream);
  const int tiles = (N + tcc::kTileM - 1) / tcc::kTileM;
  const uint8_t *s8 = static_cast<const uint8_t *>(s_img), *h8 = static_cast<const uint8_t *>(h_img), *w8 = static_cast<const uint8_t *>(workspace);
  uint8_t *o8 = static_cast<uint8_t *>(h_out_img);
  if (!g_fwd_cluster) {
    DDFA_CUDA(cudaFuncSetAttribute(tc2::gru_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::kSmemAlloc));
    int groups = kNumSMs / tc2::kSlices;
    if (groups > tiles) groups = tiles;
    tc2::gru_fwd_kernel<false><<<groups * tc2::kSlices, tc2::kThreads, tc2::kSmemAlloc, stream>>>(s8, h8, h, indptr, w8, N, h_out, o8, save_gates);
    DDFA_CHECK_LAUNCH("tc2::gru_fwd_kernel");
    return DDFA_OK;
  }
  DDFA_CUDA(cudaFuncSetAttribute(tc2::gru_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::kSmemAlloc));
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = tc2::kSlices;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(tc2::kThreads);
  cfg.dynamicSmemBytes = tc2::kSmemAlloc;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // the tile schedule assumes every cluster is resident: ask how many 4-CTA clusters fit (GPC granularity: < 148 / 4)
  static int max_clusters = 0;
  if (max_clusters == 0) {
    cfg.gridDim = dim3(kNumSMs / tc2::kSlices * tc2::kSlices);
    int n = 0;
    DDFA_CUDA(cudaOccupancyMaxActiveClusters(&n, tc2::gru_fwd_kernel<true>, &cfg));
    max_clusters = n > 0 ? n : 1;
  }
  int groups = max_clusters;
  if (groups > tiles) groups = tiles;
  cfg.gridDim = dim3(groups * tc2::kSlices);
  DDFA_CUDA(cudaLaunchKernelEx(&cfg, tc2::gru_fwd_kernel<true>, s8, h8, h, indptr, w8, (int32_t)N, h_out, o8, save_gates));
  count_launch();
  return DDFA_OK;
}

}  // namespace ddfa
