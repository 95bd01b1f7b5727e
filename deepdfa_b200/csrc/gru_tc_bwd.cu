// tcgen05 engine, backward of one GRU step (D == 128) — activation images in, TMA-fed, persistent kernels.
// Backward of the same math the forward kernel implements (gru_tc_fwd3.cu; reference: DDFA/code_gnn/models/flow_gnn/
// ggnn.py:60-63 through dgl.nn.GatedGraphConv / torch.nn.GRUCell autograd).
//
//   (1) gate_bwd_image_kernel   q_r, q_z, q_n, q_nr  <-  (dh', h, r, z, n, gh_n)      elementwise, HBM-bound
//         writes the four q matrices as activation IMAGES, the plane dh' * z (dgrad's elementwise term) and the bias
//         gradients (column sums).  The incoming gradient may be given in two parts, dh' = dh_part + A^T ds_prev: the
//         transposed edge gather of the previous step's ds is folded into this kernel's row loop (no dh' round trip
//         through HBM and one launch less per step).
//   (2) dgrad3_kernel           ds = [q_r q_z q_n] W' ;  dh = dh' * z + [q_r q_z q_nr] Whh      K = 3D
//         transposed GEMM: the weights live in tensor memory as the A operand, the q images stream through three 64 KB
//         stages as the B operand, a CTA owns all 128 columns of ds or of dh (see the comment at the kernel)
//   (3) wgrad_kernel            dW' += [q_r q_z q_n]^T s ;  dWhh += [q_r q_z q_nr]^T h            K = nodes
//         both operands are read "MN-major" straight from the images (whole 128-node tiles, three 64 KB slots); a CTA keeps
//         its [384 x 128] fp32 partial sum in TMEM over all its tiles and folds it into a private global partial at the
//         end; wgrad_reduce_kernel sums the partials once per backward pass.
// Precision: bf16x3 everywhere (hi*hi + hi*lo + lo*hi), fp32 accumulate.
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace ddfa {
namespace tc2b {
using namespace tcc;

constexpr int kThreads = 320, kEpiWarps = 8;

// =================================================================================================
// (1) gate backward -> q images, h image, bias gradients
// =================================================================================================
// Persistent: 2 CTAs per SM, every warp strides over node rows (lane = 4 columns), two rows in flight per warp (12 x 16-byte
// loads outstanding per lane); the seven column sums (bias gradients) stay in registers for the whole kernel and are
// combined once per CTA (r01s ncu: with one CTA per 64 rows the shared/global atomics of that reduction were 22% of the
// kernel and the row loop was load-latency-bound).
constexpr int kGbWarps = 8;
struct GbRow {
  float4 d, hv, rr, zz, nn, gh;
  float deg;
  int tb, te;      // this row's neighbour range in the transposed CSR (fused gather)
};
__device__ __forceinline__ float4 bf16x4_to_f4(const uint2 &p) {
  return make_float4(__uint_as_float(p.x << 16), __uint_as_float(p.x & 0xffff0000u), __uint_as_float(p.y << 16), __uint_as_float(p.y & 0xffff0000u));
}
// Two forms of the saved forward state:
//   fp32:   h = [N,128] plane, gates = four [N,128] planes (r, z, n, gh_n)                       (ddfa_gru_step_bwd_image)
//   packed: h = the activation image the forward GEMM read (hi + lo), gates_packed = [N,128] x 64-bit words (pack_gates)
//           — 64 instead of 96 bytes per lane-row                                               (ddfa_gru_step_bwd_image_v2)
__device__ __forceinline__ void gb_load(GbRow &x, const float *__restrict__ dh_out, const float *__restrict__ h,
                                        const uint8_t *__restrict__ h_img_src, const float *__restrict__ gates,
                                        const uint4 *__restrict__ gates_packed, const int32_t *__restrict__ indptr,
                                        const int32_t *__restrict__ indptr_t, size_t plane, int64_t node, int col, bool ok,
                                        uint64_t pol_saved, uint64_t pol_dh) {
  x.tb = x.te = 0;
  if (ok) {
    if (indptr_t) { x.tb = __ldcg(indptr_t + node); x.te = __ldcg(indptr_t + node + 1); }
    const size_t off = (size_t)node * kD + col;
    x.d = ldg_cg_f4_hint(dh_out + off, pol_dh);
    uint2 hh = make_uint2(0u, 0u), hl = hh;
    uint4 g0 = make_uint4(0u, 0u, 0u, 0u), g1 = g0;
    if (h) x.hv = ldg_cg_f4_hint(h + off, pol_saved);                 // saved activations: last use
    else {
      hh = __ldcg(reinterpret_cast<const uint2 *>(h_img_src + image_offset(node, col, 0)));
      hl = __ldcg(reinterpret_cast<const uint2 *>(h_img_src + image_offset(node, col, 1)));
    }
    if (gates) {
      x.rr = ldg_cg_f4_hint(gates + off, pol_saved);
      x.zz = ldg_cg_f4_hint(gates + plane + off, pol_saved);
      x.nn = ldg_cg_f4_hint(gates + 2 * plane + off, pol_saved);
      x.gh = ldg_cg_f4_hint(gates + 3 * plane + off, pol_saved);
    } else {
      g0 = __ldcg(gates_packed + (off >> 1));          // columns col, col+1: {rz, ng, rz, ng}
      g1 = __ldcg(gates_packed + (off >> 1) + 1);      // columns col+2, col+3
    }
    x.deg = (float)(__ldcg(indptr + node + 1) - __ldcg(indptr + node));
    if (!h) {
      const float4 a = bf16x4_to_f4(hh), b = bf16x4_to_f4(hl);
      x.hv = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
    if (!gates) {
      unpack_gates(make_uint2(g0.x, g0.y), x.rr.x, x.zz.x, x.nn.x, x.gh.x);
      unpack_gates(make_uint2(g0.z, g0.w), x.rr.y, x.zz.y, x.nn.y, x.gh.y);
      unpack_gates(make_uint2(g1.x, g1.y), x.rr.z, x.zz.z, x.nn.z, x.gh.z);
      unpack_gates(make_uint2(g1.z, g1.w), x.rr.w, x.zz.w, x.nn.w, x.gh.w);
    }
  }
}

__global__ void __launch_bounds__(32 * kGbWarps, 2) gate_bwd_image_kernel(const float *__restrict__ dh_out, const float *__restrict__ h,
                                                                          const uint8_t *__restrict__ h_img_src,
                                                                          const float *__restrict__ gates, const uint4 *__restrict__ gates_packed,
                                                                          const int32_t *__restrict__ indptr,
                                                                          const float *__restrict__ ds_in, const int32_t *__restrict__ indptr_t,
                                                                          const int32_t *__restrict__ indices_t,
                                                                          int32_t N, uint8_t *__restrict__ q_img, size_t img_stride,
                                                                          uint8_t *__restrict__ h_img, float *__restrict__ dhz,
                                                                          float *__restrict__ db_fold,
                                                                          float *__restrict__ db_ih, float *__restrict__ db_hh, int hints) {
  __shared__ float red[kGbWarps][7 * kD];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = lane * 4;
  pdl_launch_dependents();
  pdl_wait();
  const uint64_t pol_saved = l2_policy((hints & 2) ? 1 : 0), pol_tmp = l2_policy((hints & 4) ? 2 : 0), pol_dh = l2_policy((hints & 8) ? 1 : 0);
  const size_t plane = (size_t)N * kD;
  float4 sum[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) sum[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t Npad = ((int64_t)N + kTileM - 1) / kTileM * kTileM;
  const int64_t stride = (int64_t)gridDim.x * kGbWarps;

  auto finish = [&](const GbRow &x, int64_t node, bool ok) {
    float4 qr = make_float4(0.f, 0.f, 0.f, 0.f), qz = qr, qn = qr, qnr = qr, hv = qr;
    if (ok) {
      hv = x.hv;
      st_f4_hint(dhz + (size_t)node * kD + col, make_float4(x.d.x * x.zz.x, x.d.y * x.zz.y, x.d.z * x.zz.z, x.d.w * x.zz.w), pol_tmp);
#define BWDQ(f)                                                      \
  {                                                                  \
    const float dz_ = x.d.f * (x.hv.f - x.nn.f);                     \
    const float dn_ = x.d.f * (1.f - x.zz.f);                        \
    qn.f = dn_ * (1.f - x.nn.f * x.nn.f);                            \
    qz.f = dz_ * x.zz.f * (1.f - x.zz.f);                            \
    qr.f = qn.f * x.gh.f * x.rr.f * (1.f - x.rr.f);                  \
    qnr.f = qn.f * x.rr.f;                                           \
  }
      BWDQ(x) BWDQ(y) BWDQ(z) BWDQ(w)
#undef BWDQ
      f4_add(sum[0], qr); f4_add(sum[1], qz); f4_add(sum[2], qn); f4_add(sum[3], qnr);
      f4_fma(sum[4], x.deg, qr); f4_fma(sum[5], x.deg, qz); f4_fma(sum[6], x.deg, qn);
    }
    // images: rows N..Npad-1 are written as zeros (the weight-gradient GEMM sums over all 128 rows of a tile)
    const size_t o_hi = image_offset(node, col, 0), o_lo = image_offset(node, col, 1);
    uint2 ph, pl;
    split4(qr, ph, pl);  *reinterpret_cast<uint2 *>(q_img + 0 * img_stride + o_hi) = ph; *reinterpret_cast<uint2 *>(q_img + 0 * img_stride + o_lo) = pl;
    split4(qz, ph, pl);  *reinterpret_cast<uint2 *>(q_img + 1 * img_stride + o_hi) = ph; *reinterpret_cast<uint2 *>(q_img + 1 * img_stride + o_lo) = pl;
    split4(qn, ph, pl);  *reinterpret_cast<uint2 *>(q_img + 2 * img_stride + o_hi) = ph; *reinterpret_cast<uint2 *>(q_img + 2 * img_stride + o_lo) = pl;
    split4(qnr, ph, pl); *reinterpret_cast<uint2 *>(q_img + 3 * img_stride + o_hi) = ph; *reinterpret_cast<uint2 *>(q_img + 3 * img_stride + o_lo) = pl;
    if (h_img) {   // only when the caller did not keep the forward pass's image of h_t
      split4(hv, ph, pl); *reinterpret_cast<uint2 *>(h_img + o_hi) = ph; *reinterpret_cast<uint2 *>(h_img + o_lo) = pl;
    }
  };

  for (int64_t node = (int64_t)blockIdx.x * kGbWarps + warp; node < Npad; node += 2 * stride) {
    const int64_t node2 = node + stride;
    GbRow a, b;
    const bool ok_a = node < N, ok_b = node2 < N;
    gb_load(a, dh_out, h, h_img_src, gates, gates_packed, indptr, indptr_t, plane, node, col, ok_a, pol_saved, pol_dh);
    gb_load(b, dh_out, h, h_img_src, gates, gates_packed, indptr, indptr_t, plane, node2, col, ok_b, pol_saved, pol_dh);
    if (indptr_t) {
      // dh' += sum over the transposed-graph neighbours of ds_in: first up to 4 neighbour ids of both rows, then their
      // rows, all loads of a phase in flight together; longer lists finish in a plain loop (deterministic order)
      int ia[4], ib[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ia[k] = (a.tb + k < a.te) ? __ldcg(indices_t + a.tb + k) : -1;
        ib[k] = (b.tb + k < b.te) ? __ldcg(indices_t + b.tb + k) : -1;
      }
      float4 va[4], vb[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        va[k] = ia[k] >= 0 ? ldg_cg_f4(ds_in + (size_t)ia[k] * kD + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        vb[k] = ib[k] >= 0 ? ldg_cg_f4(ds_in + (size_t)ib[k] * kD + col) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) { f4_add(a.d, va[k]); f4_add(b.d, vb[k]); }
      for (int j = a.tb + 4; j < a.te; ++j) f4_add(a.d, ldg_cg_f4(ds_in + (size_t)__ldcg(indices_t + j) * kD + col));
      for (int j = b.tb + 4; j < b.te; ++j) f4_add(b.d, ldg_cg_f4(ds_in + (size_t)__ldcg(indices_t + j) * kD + col));
    }
    finish(a, node, ok_a);
    if (node2 < Npad) finish(b, node2, ok_b);
  }
#pragma unroll
  for (int i = 0; i < 7; ++i) *reinterpret_cast<float4 *>(&red[warp][i * kD + col]) = sum[i];
  __syncthreads();
  for (int i = threadIdx.x; i < 7 * kD; i += 32 * kGbWarps) {
    float v_ = 0.f;
#pragma unroll
    for (int w = 0; w < kGbWarps; ++w) v_ += red[w][i];
    const int which = i >> 7, c_ = i & 127;
    // 0:S(q_r) 1:S(q_z) 2:S(q_n) 3:S(q_nr) 4:S(deg q_r) 5:S(deg q_z) 6:S(deg q_n)
    if (which == 0) { atomicAdd(db_ih + c_, v_); atomicAdd(db_hh + c_, v_); }
    else if (which == 1) { atomicAdd(db_ih + kD + c_, v_); atomicAdd(db_hh + kD + c_, v_); }
    else if (which == 2) atomicAdd(db_ih + 2 * kD + c_, v_);
    else if (which == 3) atomicAdd(db_hh + 2 * kD + c_, v_);
    else atomicAdd(db_fold + (which - 4) * kD + c_, v_);
  }
}

// -------------------------------------------------------------------------------------------------
// (1') gate backward, TMA-staged (packed saved state).  The register-path kernel above is load-latency-bound: ncu on C1
// (profiles/r03i) shows 170 us for 731 MB = 4.3 TB/s, DRAM 52 %, 85 % of the stall samples on long_scoreboard — 16 warps per SM
// with ~128 bytes of streaming loads in flight per lane do not cover the loaded DRAM latency, and the register file (128 x 512)
// admits no more.  Here the three STREAMING operands of a row block — dh (fp32), the packed gates and h (image pieces, or fp32 for
// step 0) — are contiguous in memory, so a producer warp moves them with TMA bulk copies into a three-stage shared-memory ring
// (3 x 64 KB in flight per SM, no registers), and 16 consumer warps read them with LDS; only the gathered ds rows of the folded
// transposed edge gather (data-dependent addresses) remain register loads, and they hit L2 (dgrad wrote them a kernel ago).
// One CTA per SM, block of 32 consecutive rows per stage, warp w owns rows w and w + 16 of the block, lane = 4 columns as before:
// same arithmetic, same summation order of the gather, same outputs.
// -------------------------------------------------------------------------------------------------
constexpr int kGtRows = 32, kGtStages = 3, kGtConsumers = 16, kGtPre = 2;
constexpr int kGtOffD = 0, kGtOffG = kGtRows * kD * 4, kGtOffH = kGtOffG + kGtRows * kD * 8;          // 16 KB | 32 KB | 16 KB
constexpr int kGtStageBytes = kGtOffH + kGtRows * kD * 4;                                              // 64 KB
constexpr int kGtOffBar = kGtStages * kGtStageBytes;
constexpr int kGtSmem = kGtOffBar + 2 * kGtStages * 8 + 16;
constexpr int kGtThreads = 32 * kGtConsumers;

// HF32: h operand as fp32 rows (step 0: h_0 = x) or the activation image.  CSRP: the CSR scalars / neighbour ids of the folded
// gather pipelined across iterations, one value per lane (DDFA_TUNE_GATE_BWD_TMA = 2, default; 1 = fetched inside the iteration).
template <bool HF32, bool CSRP>
__global__ void __launch_bounds__(kGtThreads, 1) gate_bwd_tma_kernel(const float *__restrict__ dh_out, const float *__restrict__ h,
                                                                     const uint8_t *__restrict__ h_img_src, const uint2 *__restrict__ gates_packed,
                                                                     const int32_t *__restrict__ indptr, const float *__restrict__ ds_in,
                                                                     const int32_t *__restrict__ indptr_t, const int32_t *__restrict__ indices_t,
                                                                     int32_t N, uint8_t *__restrict__ q_img, size_t img_stride,
                                                                     float *__restrict__ dhz, float *__restrict__ db_fold, float *__restrict__ db_ih,
                                                                     float *__restrict__ db_hh, int hints) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + kGtOffBar;
  auto full = [&](int i) { return bar0 + 8u * i; };
  auto empty = [&](int i) { return bar0 + 8u * (kGtStages + i); };
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t Npad = ((int64_t)N + kTileM - 1) / kTileM * kTileM;
  const int num_blocks = (int)(Npad / kGtRows);
  const int my_blocks = (num_blocks > (int)blockIdx.x) ? (num_blocks - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kGtStages; ++i) { mbar_init(full(i), 1); mbar_init(empty(i), kGtConsumers); }
    mbar_fence_init();
  }
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();
  float4 sum[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) sum[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  // the copies of block k into its stage: issued by one lane of warp 0 — ahead of the loop for the first kGtStages blocks, then
  // each time warp 0 has finished a block (a dedicated producer warp would make 17 warps = 640 allocated threads = 96 registers;
  // 16 warps get 128 and the kernel does not spill)
  auto fill = [&](int k) {
    const uint64_t pol_saved = l2_policy((hints & 2) ? 1 : 0), pol_dh = l2_policy((hints & 8) ? 1 : 0);
    const int stage = k % kGtStages;
    const int64_t r0 = (int64_t)(blockIdx.x + (int64_t)k * gridDim.x) * kGtRows;
    const int valid = (int)max((int64_t)0, min((int64_t)kGtRows, (int64_t)N - r0));      // rows that exist in the [N, ...] arrays
    const uint32_t dst = sbase + stage * kGtStageBytes;
    const uint32_t bytes = (uint32_t)valid * (kD * 4 + kD * 8) + (HF32 ? (uint32_t)valid * kD * 4 : (uint32_t)kGtRows * kD * 4);
    mbar_arrive_expect_tx(full(stage), bytes);
    if (valid > 0) {
      bulk_g2s_hint(dst + kGtOffD, dh_out + r0 * kD, (uint32_t)valid * kD * 4, full(stage), pol_dh);
      bulk_g2s_hint(dst + kGtOffG, gates_packed + r0 * kD, (uint32_t)valid * kD * 8, full(stage), pol_saved);
    }
    if (HF32) {
      if (valid > 0) bulk_g2s_hint(dst + kGtOffH, h + r0 * kD, (uint32_t)valid * kD * 4, full(stage), pol_saved);
    } else {      // the block's 32 rows of each of the four [128 x 64] bf16 chunks: 4 KB pieces (the image is padded to whole tiles)
      const uint8_t *tile = h_img_src + (size_t)(r0 >> 7) * kImageTileBytes + (size_t)(r0 & 127) * 128;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) bulk_g2s_hint(dst + kGtOffH + ch * (kGtRows * 128), tile + (size_t)ch * kChunkBytes, kGtRows * 128, full(stage), pol_saved);
    }
  };
  if (warp == 0 && elect_one())
    for (int k = 0; k < kGtStages && k < my_blocks; ++k) fill(k);
  __syncwarp();
  {
    // ===== consumers =====
    const int col = lane * 4;
    const uint64_t pol_tmp = l2_policy((hints & 4) ? 2 : 0);
    // The CSR data of the warp's two rows — indptr (in-degree), indptr_t and the first kGtPre transposed neighbour ids — is a chain of
    // dependent global loads (indptr_t -> indices_t -> ds row); fetched inside the iteration that uses it, the chain was 59 % of the
    // kernel's stall samples (profiles/r03r: long_scoreboard on these lines).  It is warp-uniform, so it is kept ONE VALUE PER LANE
    // and pipelined across iterations: lanes 0-3 hold indptr[node_r + {0,1}], lanes 4-7 indptr_t[node_r + {0,1}] (r = (lane >> 1) & 1)
    // of a block, lanes 8-11 its ids id[r][q] (q = lane & 1); iteration k requests the scalars of block k + 2 and the ids of block
    // k + 1 (from the scalars that arrived during iteration k - 1) and broadcasts block k's values by shuffle — only the ds rows
    // themselves are still requested in the iteration that adds them.
    static_assert(kGtPre == 2, "lane slots below assume two prefetched neighbours per row");
    auto load_scalars = [&](int kk) -> int {
      int v = 0;
      if (kk < my_blocks && lane < 8) {
        const int64_t nd = (int64_t)(blockIdx.x + (int64_t)kk * gridDim.x) * kGtRows + warp + 16 * ((lane >> 1) & 1);
        const int32_t *base = lane < 4 ? indptr : indptr_t;
        if (nd < N && base) v = __ldcg(base + nd + (lane & 1));
      }
      return v;
    };
    auto load_ids = [&](int sc) -> int {
      const int r_ = (lane >> 1) & 1, q_ = lane & 1;
      const int tb_ = __shfl_sync(0xffffffffu, sc, 4 + 2 * r_), te_ = __shfl_sync(0xffffffffu, sc, 5 + 2 * r_);
      int v = -1;
      if (lane >= 8 && lane < 12 && tb_ + q_ < te_) v = __ldcg(indices_t + tb_ + q_);
      return v;
    };
    int sc_cur = 0, sc_next = 0, id_cur = -1;
    if constexpr (CSRP) {
      sc_cur = load_scalars(0); sc_next = load_scalars(1);
      id_cur = load_ids(sc_cur);
    }
    for (int k = 0; k < my_blocks; ++k) {
      const int stage = k % kGtStages, use = k / kGtStages;
      const int64_t r0 = (int64_t)(blockIdx.x + (int64_t)k * gridDim.x) * kGtRows;
      // the two rows of this warp and their CSR data
      int64_t node[2];
      bool ok[2];
      int tb[2], te[2], ip0[2], ip1[2];
      int id[2][kGtPre];      // the first kGtPre neighbours of both rows (a CFG node has ~2 in-edges); longer lists finish below
      if constexpr (CSRP) {
        const int sc_next2 = load_scalars(k + 2);
        const int id_next = load_ids(sc_next);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          node[r] = r0 + warp + 16 * r;
          ok[r] = node[r] < N;
          ip0[r] = __shfl_sync(0xffffffffu, sc_cur, 2 * r); ip1[r] = __shfl_sync(0xffffffffu, sc_cur, 2 * r + 1);
          tb[r] = __shfl_sync(0xffffffffu, sc_cur, 4 + 2 * r); te[r] = __shfl_sync(0xffffffffu, sc_cur, 5 + 2 * r);
#pragma unroll
          for (int q = 0; q < kGtPre; ++q) id[r][q] = __shfl_sync(0xffffffffu, id_cur, 8 + 2 * r + q);
        }
        sc_cur = sc_next; sc_next = sc_next2; id_cur = id_next;
      } else {      // everything requested inside the iteration (global, independent of the staged operands: before the wait)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          node[r] = r0 + warp + 16 * r;
          ok[r] = node[r] < N;
          tb[r] = te[r] = ip0[r] = ip1[r] = 0;
          if (ok[r]) {
            ip0[r] = __ldcg(indptr + node[r]); ip1[r] = __ldcg(indptr + node[r] + 1);
            if (indptr_t) { tb[r] = __ldcg(indptr_t + node[r]); te[r] = __ldcg(indptr_t + node[r] + 1); }
          }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int q = 0; q < kGtPre; ++q) id[r][q] = (tb[r] + q < te[r]) ? __ldcg(indices_t + tb[r] + q) : -1;
      }
      float4 gv[2][kGtPre];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int q = 0; q < kGtPre; ++q) gv[r][q] = id[r][q] >= 0 ? ldg_cg_f4(ds_in + (size_t)id[r][q] * kD + col) : make_float4(0.f, 0.f, 0.f, 0.f);
      mbar_wait(full(stage), use & 1);
      const uint8_t *st = smem + stage * kGtStageBytes;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int row = warp + 16 * r;                   // row inside the block
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f), hv = d, rr = d, zz = d, nn = d, gh = d;
        if (ok[r]) {
          d = *reinterpret_cast<const float4 *>(st + kGtOffD + row * (kD * 4) + col * 4);
          const uint4 g0 = *reinterpret_cast<const uint4 *>(st + kGtOffG + row * (kD * 8) + col * 8);
          const uint4 g1 = *reinterpret_cast<const uint4 *>(st + kGtOffG + row * (kD * 8) + col * 8 + 16);
          unpack_gates(make_uint2(g0.x, g0.y), rr.x, zz.x, nn.x, gh.x);
          unpack_gates(make_uint2(g0.z, g0.w), rr.y, zz.y, nn.y, gh.y);
          unpack_gates(make_uint2(g1.x, g1.y), rr.z, zz.z, nn.z, gh.z);
          unpack_gates(make_uint2(g1.z, g1.w), rr.w, zz.w, nn.w, gh.w);
          if (HF32) {
            hv = *reinterpret_cast<const float4 *>(st + kGtOffH + row * (kD * 4) + col * 4);
          } else {      // piece [v][kb = col / 64] of 32 rows x 128 B; 16-byte units swizzled by (global row) & 7 == row & 7 (blocks start at multiples of 32)
            const uint32_t off = (uint32_t)((col >> 6) * (kGtRows * 128) + row * 128 + (((((col & 63) >> 3) ^ (row & 7)) & 7) << 4) + (col & 7) * 2);
            const uint2 hh = *reinterpret_cast<const uint2 *>(st + kGtOffH + off);
            const uint2 hl = *reinterpret_cast<const uint2 *>(st + kGtOffH + 2 * (kGtRows * 128) + off);
            const float4 a = bf16x4_to_f4(hh), b = bf16x4_to_f4(hl);
            hv = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
          }
#pragma unroll
          for (int q = 0; q < kGtPre; ++q) f4_add(d, gv[r][q]);
          for (int j = tb[r] + kGtPre; j < te[r]; ++j) f4_add(d, ldg_cg_f4(ds_in + (size_t)__ldcg(indices_t + j) * kD + col));
        }
        float4 qr = make_float4(0.f, 0.f, 0.f, 0.f), qz = qr, qn = qr, qnr = qr;
        if (ok[r]) {
          const float deg = (float)(ip1[r] - ip0[r]);
          st_f4_hint(dhz + (size_t)node[r] * kD + col, make_float4(d.x * zz.x, d.y * zz.y, d.z * zz.z, d.w * zz.w), pol_tmp);
#define BWDQ2(f)                                         \
  {                                                      \
    const float dz_ = d.f * (hv.f - nn.f);               \
    const float dn_ = d.f * (1.f - zz.f);                \
    qn.f = dn_ * (1.f - nn.f * nn.f);                    \
    qz.f = dz_ * zz.f * (1.f - zz.f);                    \
    qr.f = qn.f * gh.f * rr.f * (1.f - rr.f);            \
    qnr.f = qn.f * rr.f;                                 \
  }
          BWDQ2(x) BWDQ2(y) BWDQ2(z) BWDQ2(w)
#undef BWDQ2
          f4_add(sum[0], qr); f4_add(sum[1], qz); f4_add(sum[2], qn); f4_add(sum[3], qnr);
          f4_fma(sum[4], deg, qr); f4_fma(sum[5], deg, qz); f4_fma(sum[6], deg, qn);
        }
        if (node[r] < Npad) {      // rows N .. Npad-1 are written as zeros (the weight-gradient GEMM sums over all 128 rows of a tile)
          const size_t o_hi = image_offset(node[r], col, 0), o_lo = image_offset(node[r], col, 1);
          uint2 ph, pl;
          split4(qr, ph, pl);  *reinterpret_cast<uint2 *>(q_img + 0 * img_stride + o_hi) = ph; *reinterpret_cast<uint2 *>(q_img + 0 * img_stride + o_lo) = pl;
          split4(qz, ph, pl);  *reinterpret_cast<uint2 *>(q_img + 1 * img_stride + o_hi) = ph; *reinterpret_cast<uint2 *>(q_img + 1 * img_stride + o_lo) = pl;
          split4(qn, ph, pl);  *reinterpret_cast<uint2 *>(q_img + 2 * img_stride + o_hi) = ph; *reinterpret_cast<uint2 *>(q_img + 2 * img_stride + o_lo) = pl;
          split4(qnr, ph, pl); *reinterpret_cast<uint2 *>(q_img + 3 * img_stride + o_hi) = ph; *reinterpret_cast<uint2 *>(q_img + 3 * img_stride + o_lo) = pl;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty(stage));      // this warp has read its rows of the stage
      if (warp == 0 && k + kGtStages < my_blocks) {  // refill the stage with block k + kGtStages once all 16 warps have released it
        if (elect_one()) {
          mbar_wait(empty(stage), use & 1);
          fill(k + kGtStages);
        }
        __syncwarp();
      }
    }
  }
  // bias gradients: column sums of the 16 consumer warps, through the (now idle) first stage
  __syncthreads();
  float *red = reinterpret_cast<float *>(smem);
  {
#pragma unroll
    for (int i = 0; i < 7; ++i) *reinterpret_cast<float4 *>(&red[(warp * 7 + i) * kD + lane * 4]) = sum[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 7 * kD; i += kGtThreads) {
    float v_ = 0.f;
#pragma unroll
    for (int w = 0; w < kGtConsumers; ++w) v_ += red[(w * 7) * kD + i];
    const int which = i >> 7, c_ = i & 127;
    if (which == 0) { atomicAdd(db_ih + c_, v_); atomicAdd(db_hh + c_, v_); }
    else if (which == 1) { atomicAdd(db_ih + kD + c_, v_); atomicAdd(db_hh + kD + c_, v_); }
    else if (which == 2) atomicAdd(db_ih + 2 * kD + c_, v_);
    else if (which == 3) atomicAdd(db_hh + 2 * kD + c_, v_);
    else atomicAdd(db_fold + (which - 4) * kD + c_, v_);
  }
}

// =================================================================================================
// (2) dgrad, weight-in-TMEM orientation ("dgrad3")
// =================================================================================================
// The in-kernel timeline of its predecessor (weight slices in shared memory; profiles/r01l_trace_dgrad.log) showed the operand feed as the bound: 96 KB of
// shared memory hold the weight slice, only 2 x 32 KB are left for activations, and with ~1 us per copy in flight that
// is ~35 GB/s per SM while every CTA has to pull 256 KB per tile (each q tile is read by four slice CTAs).
// Here the GEMM is transposed:  D^T[col, node] = W^T[col, K] * Q[node, K]^T
//   * A = the weights, held in TENSOR MEMORY for the life of the CTA (lane = output column, two bf16 per 32-bit column:
//     K = 3 x 128 -> 192 columns hi + 192 columns lo), loaded once with tcgen05.st from a packed global array;
//   * B = the q images (K-major SWIZZLE_128B, N = nodes), streamed through THREE 64 KB stages — all of shared memory
//     is pipeline now (copy_bench2: 64 KB x 3 stages feeds ~130 GB/s per SM);
//   * a CTA owns all 128 columns of one output (role 0: ds = [q_r q_z q_n] W', role 1: dh = [q_r q_z q_nr] Whh), so a q
//     tile is read by 2 CTAs instead of 4 and 192 KB are pulled per 128 x 128 outputs (was 256 KB per 128 x 64);
//   * D = two 64-node accumulator halves (64 TMEM columns each) so the epilogue of one half overlaps the MMAs of the other;
//   * the epilogue thread holds one output column for 16 nodes per tcgen05.ld: a warp-wide store covers 32 consecutive
//     columns of one node = one full 128-byte line, no shared-memory staging.
constexpr int kD3Stages = 3;
constexpr int kD3StageBytes = kImageTileBytes;               // one q-matrix tile: [hi|lo][kb0|kb1], 64 KB
constexpr int kD3WColsHalf = 192;                            // K = 384 bf16 -> 192 packed columns per variant
constexpr int kD3AccCol = 2 * kD3WColsHalf;                  // accumulators start at TMEM column 384
constexpr int kD3OffBar = kD3Stages * kD3StageBytes;         // 192 KB
constexpr int kD3NumBars = 2 * kD3Stages + 4 + 1;            // a_full, a_empty, acc_full[2], acc_empty[2], w_ready
constexpr int kD3OffTmemPtr = kD3OffBar + kD3NumBars * 8;
constexpr int kD3SmemAlloc = kD3OffTmemPtr + 16 + 1024;
constexpr int kD3Chunks = 2 * kD3WColsHalf / 16;             // 24 chunks of 16 TMEM columns
constexpr size_t kD3PackedBytes = (size_t)2 * kD3Chunks * 128 * 64;   // [role][chunk][lane][16 words] = 384 KB

// packed[role][chunk][lane j][i]: TMEM column chunk*16+i of lane j = bf16 pair (W[kk][j], W[kk+1][j]), hi for columns
// 0..191, lo for 192..383, kk = 2 * (column mod 192); W = W' (role 0) or Whh (role 1), both [3D x D] row-major.
__global__ void dgrad3_pack_kernel(const float *__restrict__ w_fold, const float *__restrict__ w_hh, uint32_t *__restrict__ packed) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * kD3Chunks * 128) return;
  const int lane = idx % 128, chunk = (idx / 128) % kD3Chunks, role = idx / (128 * kD3Chunks);
  const float *W = role == 0 ? w_fold : w_hh;
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int col = chunk * 16 + i;
    const int v = col >= kD3WColsHalf ? 1 : 0;
    const int kk = (col - v * kD3WColsHalf) * 2;
    __nv_bfloat16 h0, l0, h1, l1;
    split_bf16(W[(size_t)kk * kD + lane], h0, l0);
    split_bf16(W[(size_t)(kk + 1) * kD + lane], h1, l1);
    w[i] = v ? ((uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16))
             : ((uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16));
  }
  uint4 *dst = reinterpret_cast<uint4 *>(packed + (size_t)idx * 16);
#pragma unroll
  for (int i = 0; i < 4; ++i) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}

__global__ void __launch_bounds__(kThreads, 1) dgrad3_kernel(const uint8_t *__restrict__ q_img, size_t img_stride,
                                                             const float *__restrict__ dhz,
                                                             const uint32_t *__restrict__ packed3, int32_t N,
                                                             float *__restrict__ ds, float *__restrict__ dh, int hints) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + kD3OffBar;
  auto a_full = [&](int i) { return bar0 + 8u * i; };
  auto a_empty = [&](int i) { return bar0 + 8u * (kD3Stages + i); };
  auto acc_full = [&](int i) { return bar0 + 8u * (2 * kD3Stages + i); };
  auto acc_empty = [&](int i) { return bar0 + 8u * (2 * kD3Stages + 2 + i); };
  const uint32_t w_ready = bar0 + 8u * (2 * kD3Stages + 4);
  volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(smem + kD3OffTmemPtr);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int role = blockIdx.x & 1;
  const int group = blockIdx.x >> 1, num_groups = gridDim.x >> 1;
  const int num_tiles = (N + kTileM - 1) / kTileM;
  const int my_tiles = (num_tiles > group) ? (num_tiles - 1 - group) / num_groups + 1 : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kD3Stages; ++i) { mbar_init(a_full(i), 1); mbar_init(a_empty(i), 1); }
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
  const int tron = (g_trace_on == 1);
  if (threadIdx.x == 0) trace_stamp(tron, 0, 0);
  pdl_launch_dependents();

  if (warp == 0) {
    // ===== producer: per tile the three q matrices of this role, one 64 KB copy each =====
    if (elect_one()) {
      pdl_wait();      // the q images come from gate_bwd_image_kernel, the previous kernel of the chain
      const uint64_t pol_q = l2_policy((hints & 64) ? 1 : 0);
      int cc = 0;
      for (int k = 0; k < my_tiles; ++k) {
        const int tile = num_tiles - 1 - (group + k * num_groups);   // back to front: the q tiles written last are still in L2
        for (int g = 0; g < 3; ++g, ++cc) {
          const int m = g < 2 ? g : (role == 0 ? 2 : 3);      // q_r, q_z, then q_n (ds) or q_nr (dh)
          const int stage = cc % kD3Stages, use = cc / kD3Stages;
          if (use > 0) mbar_wait(a_empty(stage), (use - 1) & 1);
          if (g == 0) trace_stamp(tron, k, 1);
          mbar_arrive_expect_tx(a_full(stage), kD3StageBytes);
          bulk_g2s_hint(sbase + stage * kD3StageBytes, q_img + (size_t)m * img_stride + (size_t)tile * kImageTileBytes, kD3StageBytes,
                        a_full(stage), pol_q);
          if (g == 2) trace_stamp(tron, k, 2);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (my_tiles > 0 && elect_one()) {
      constexpr uint32_t kIdesc64 = make_idesc(64);
      mbar_wait(w_ready, 0);
      tc_fence_after();
      int cc = 0;
      for (int k = 0; k < my_tiles; ++k) {
        for (int g = 0; g < 3; ++g, ++cc) {
          const int stage = cc % kD3Stages, use = cc / kD3Stages;
          mbar_wait(a_full(stage), use & 1);
          tc_fence_after();
          if (g == 0) trace_stamp(tron, k, 4);
          if (g == 2) trace_stamp(tron, k, 5);
          for (int half = 0; half < 2; ++half) {
            if (g == 0 && k > 0) {
              mbar_wait(acc_empty(half), (k - 1) & 1);
              tc_fence_after();
              if (half == 0) trace_stamp(tron, k, 3);
            }
            const uint32_t d_addr = tmem_base + (uint32_t)(kD3AccCol + half * 64);
            // nodes 64*half.. of each 16 KB chunk: [hi kb0 | hi kb1 | lo kb0 | lo kb1]
            const uint64_t b_base = make_desc(sbase + stage * kD3StageBytes + (uint32_t)half * 8192u);
            const uint32_t a_base = tmem_base + (uint32_t)(g * 64);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                const uint32_t kk2 = (uint32_t)(kb * 32 + k4 * 8);                           // packed weight column of this K step
                const uint64_t b_hi = desc_advance(b_base, (uint32_t)kb * kChunkBytes + k4 * 32);
                const uint64_t b_lo = desc_advance(b_base, (uint32_t)(2 + kb) * kChunkBytes + k4 * 32);
                const uint32_t first = (g == 0 && kb == 0 && k4 == 0) ? 0u : 1u;
                umma_f16_ts(d_addr, a_base + kk2, b_hi, kIdesc64, first);                       // w_hi q_hi
                umma_f16_ts(d_addr, a_base + kD3WColsHalf + kk2, b_hi, kIdesc64, 1u);           // w_lo q_hi
                umma_f16_ts(d_addr, a_base + kk2, b_lo, kIdesc64, 1u);                          // w_hi q_lo
              }
            }
            if (g == 2) umma_commit(acc_full(half));
          }
          umma_commit(a_empty(stage));
        }
        trace_stamp(tron, k, 6);
      }
    }
  } else {
    // ===== weights -> tensor memory, then the epilogue =====
    const int q = warp & 3;                  // TMEM lane quarter: output columns 32q .. 32q+31
    const int e = (warp - 2) >> 2;           // which 32 nodes of a 64-node half (and which 12 weight chunks)
    const int col = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    if (my_tiles > 0) {
      const uint4 *src = reinterpret_cast<const uint4 *>(packed3) + ((size_t)role * kD3Chunks * 128 + (size_t)col) * 4;
#pragma unroll 4
      for (int c = 0; c < kD3Chunks / 2; ++c) {
        const int chunk = e * (kD3Chunks / 2) + c;
        const uint4 *p = src + (size_t)chunk * 128 * 4;
        const uint4 x0 = __ldcg(p), x1 = __ldcg(p + 1), x2 = __ldcg(p + 2), x3 = __ldcg(p + 3);   // L2 loads: PDL rules, common.cuh
        const uint32_t w[16] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w, x2.x, x2.y, x2.z, x2.w, x3.x, x3.y, x3.z, x3.w};
        tmem_st16(lane_addr + (uint32_t)(chunk * 16), w);
      }
      tmem_st_wait();
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(w_ready);
    pdl_wait();

    float *out = role == 0 ? ds : dh;
    const uint64_t pol_tmp = l2_policy((hints & 4) ? 2 : 0);       // ds / dh die after the next kernel has read them
    const uint64_t pol_dhz = l2_policy((hints & 8) ? 1 : 0);       // last read of dh'z
    const bool tr = (warp == 2 && lane == 0);
    for (int k = 0; k < my_tiles; ++k) {
      const int tile = num_tiles - 1 - (group + k * num_groups);
      if (tr) trace_stamp(tron, k, 7);
      for (int half = 0; half < 2; ++half) {
        const int64_t node0 = (int64_t)tile * kTileM + half * 64 + e * 32;
        int rows_valid = (int)((int64_t)N - node0);
        rows_valid = rows_valid < 0 ? 0 : (rows_valid > 32 ? 32 : rows_valid);
        float dv[32];
        if (role == 1) {       // dh = acc + (dh' * z) : fetch the elementwise term (written by gate_bwd) while the MMAs run
#pragma unroll
          for (int i = 0; i < 32; ++i) dv[i] = (i < rows_valid) ? ldg_cg_f32_hint(dhz + (node0 + i) * kD + col, pol_dhz) : 0.f;
        }
        mbar_wait(acc_full(half), k & 1);
        tc_fence_after();
        if (tr && half == 0) trace_stamp(tron, k, 8);
        const uint32_t taddr = lane_addr + (uint32_t)(kD3AccCol + half * 64 + e * 32);
        float v0[16], v1[16];
        tmem_ld16(taddr, v0);
        tmem_ld16(taddr + 16, v1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(half));
        if (tr && half == 0) trace_stamp(tron, k, 9);
        float *o = out + node0 * kD + col;
        if (role == 1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            v0[i] += dv[i];
            v1[i] += dv[16 + i];
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (i < rows_valid) st_f32_hint(o + (size_t)i * kD, v0[i], pol_tmp);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (16 + i < rows_valid) st_f32_hint(o + (size_t)(16 + i) * kD, v1[i], pol_tmp);
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

// =================================================================================================
// (3) wgrad
// =================================================================================================
// Operands are whole 128-node image tiles (64 KB, ONE bulk copy each — see the copy-size note in gru_tc_fwd.cu),
// three 64 KB slots.  Per tile the operand sequence is B_t, A_0, A_1, A_2 (B_t = s or h tile, A_g = q tiles);
// B_t must stay until A_2 is consumed, so the slots rotate:  slot(B_i) = (-i) mod 3, A_0 -> slot(B)+1, A_1 ->
// slot(B)+2, A_2 -> the slot A_0 just released.  Both operands are read MN-major (K = nodes).
constexpr int kWgSlotBytes = kImageTileBytes;             // 64 KB
constexpr int kWgSlots = 3;
constexpr int kWgOffBar = kWgSlots * kWgSlotBytes;        // 192 KB
constexpr int kWgNumBars = 2 * kWgSlots + 1;
constexpr int kWgOffTmemPtr = kWgOffBar + kWgNumBars * 8;
constexpr int kWgSmemAlloc = kWgOffTmemPtr + 16 + 1024;
constexpr size_t kWgPartialFloats = (size_t)3 * kD * kD;  // one CTA's [384 x 128] partial sum

__device__ __forceinline__ int wg_slot(int tile_i, int w) {   // w: 0 = B, 1..3 = A_0..A_2
  const int sb = (3 - tile_i % 3) % 3;
  return w == 0 ? sb : (w == 2 ? (sb + 2) % 3 : (sb + 1) % 3);
}

// grid = (ctas, 2): blockIdx.y = role: 0: A in {q_r,q_z,q_n}, B = s image -> dW' ; 1: A in {q_r,q_z,q_nr}, B = h image -> dWhh.
// partial: [2][ctas][384*128] fp32, private per CTA: accumulate (first == 0) or overwrite (first != 0); the sum over CTAs
// is taken once per backward pass by wgrad_reduce_kernel (no atomics on the hot path).
// One launch may cover several time steps (K = steps x nodes): the q images of step t are at q_img + t * step_stride, its
// B operands are batch.s_img[t] / batch.h_img[t].  Batching all T steps of a backward pass into one launch removes T-1
// epilogues (read-modify-write of the 58 MB of private partial sums), T-1 pipeline ramps and T-1 launches.
constexpr int kWgMaxSteps = 16;
struct WgBatch {
  const uint8_t *s_img[kWgMaxSteps];
  const uint8_t *h_img[kWgMaxSteps];
  int32_t steps;
};
__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(const uint8_t *__restrict__ q_img, size_t img_stride, size_t step_stride,
                                                            const __grid_constant__ WgBatch batch,
                                                            int32_t N, float *__restrict__ partial, int first, int hints) {
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
  const int num_tiles = (N + kTileM - 1) / kTileM;           // per time step
  const int all_tiles = num_tiles * batch.steps;
  const int my_tiles = (all_tiles > (int)blockIdx.x) ? (all_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgSlots; ++i) { mbar_init(full_bar(i), 1); mbar_init(empty_bar(i), 1); }
    mbar_init(acc_bar, 1);
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
  const int tron = (g_trace_on == 2) && blockIdx.y == 0;   // timeline of the role-0 CTAs (ddfa_debug_set key 2, value 2)
  if (threadIdx.x == 0) trace_stamp(tron, 0, 0);

  if (warp == 0) {
    if (elect_one()) {
      const uint64_t pol_wg = l2_policy((hints & 32) ? 1 : 0);     // every operand of the batched launch is read once
      int uses[kWgSlots] = {0, 0, 0};
      for (int i = 0; i < my_tiles; ++i) {
        const int idx = (int)blockIdx.x + i * (int)gridDim.x;
        const int t = idx / num_tiles, tile = idx - t * num_tiles;
        for (int w = 0; w < 4; ++w) {
          const int slot = wg_slot(i, w);
          if (uses[slot] > 0) mbar_wait(empty_bar(slot), (uses[slot] - 1) & 1);
          ++uses[slot];
          if (w == 0) trace_stamp(tron, i, 1);
          if (w == 3) trace_stamp(tron, i, 2);
          const uint8_t *src;
          if (w == 0) src = (role == 0) ? batch.s_img[t] : batch.h_img[t];
          else {
            const int pl = (w == 3) ? (role == 0 ? 2 : 3) : (w - 1);
            src = q_img + (size_t)t * step_stride + (size_t)pl * img_stride;
          }
          mbar_arrive_expect_tx(full_bar(slot), kWgSlotBytes);
          bulk_g2s_hint(sbase + slot * kWgSlotBytes, src + (size_t)tile * kImageTileBytes, kWgSlotBytes, full_bar(slot), pol_wg);
        }
      }
    }
  } else if (warp == 1) {
    if (my_tiles > 0 && elect_one()) {
      constexpr uint32_t kIdescMN = make_idesc(128, true, true);
      int uses[kWgSlots] = {0, 0, 0};
      for (int i = 0; i < my_tiles; ++i) {
        const int slot_b = wg_slot(i, 0);
        mbar_wait(full_bar(slot_b), uses[slot_b] & 1);
        ++uses[slot_b];
        trace_stamp(tron, i, 3);
        for (int g = 0; g < 3; ++g) {
          const int slot_a = wg_slot(i, 1 + g);
          mbar_wait(full_bar(slot_a), uses[slot_a] & 1);
          ++uses[slot_a];
          tc_fence_after();
          if (g == 0) trace_stamp(tron, i, 4);
          if (g == 2) trace_stamp(tron, i, 5);
          const uint32_t a0 = sbase + slot_a * kWgSlotBytes, b0 = sbase + slot_b * kWgSlotBytes;
          const uint32_t d_addr = tmem_base + (uint32_t)g * 128u;
#pragma unroll
          for (int k16 = 0; k16 < kTileM / 16; ++k16) {
            const uint32_t koff = (uint32_t)k16 * 2048u;     // 16 nodes = two 8-node groups of 1024 B
            const uint32_t vs = 2 * kChunkBytes;             // hi -> lo variant ([v][kb] chunks of 16 KB)
            const uint32_t acc = (i > 0 || k16 > 0) ? 1u : 0u;
            umma_f16(d_addr, make_desc_mn(a0 + koff, kChunkBytes), make_desc_mn(b0 + koff, kChunkBytes), kIdescMN, acc);       // a_hi b_hi
            umma_f16(d_addr, make_desc_mn(a0 + vs + koff, kChunkBytes), make_desc_mn(b0 + koff, kChunkBytes), kIdescMN, 1u);   // a_lo b_hi
            umma_f16(d_addr, make_desc_mn(a0 + koff, kChunkBytes), make_desc_mn(b0 + vs + koff, kChunkBytes), kIdescMN, 1u);   // a_hi b_lo
          }
          umma_commit(empty_bar(slot_a));
        }
        umma_commit(empty_bar(slot_b));
        trace_stamp(tron, i, 6);
      }
      umma_commit(acc_bar);
    }
  } else {
    // epilogue: this CTA's private partial sums.  Thread = one of the 128 rows of a gate block, this warp's 64 columns;
    // the read-modify-write of the partial slot goes through a warp-private staging tile [32 rows][68 floats] carved out
    // of the (now idle) operand ring, so global accesses are 256-byte row segments instead of 16-byte pieces.
    const int lw = warp - 2, qd = warp & 3, chalf = lw >> 2;
    constexpr int kLd = 68;
    float *dst0 = partial + ((size_t)role * gridDim.x + blockIdx.x) * kWgPartialFloats;
    if (my_tiles > 0) {
      mbar_wait(acc_bar, 0);       // every MMA (and therefore every read of the ring) has completed
      tc_fence_after();
    }
    if (warp == 2 && lane == 0) trace_stamp(tron, 0, 8);
    float *stg = reinterpret_cast<float *>(smem) + (size_t)lw * 32 * kLd;
#pragma unroll 1
    for (int g = 0; g < 3; ++g) {
      // rows 32*qd .. +31 of gate block g, columns 64*chalf .. +63: lane = (row parity, 16-byte piece).  The old partial
      // sums are requested first — all 16 loads in flight while the accumulator block is staged (the loop used to expose
      // one memory round trip per 4 rows: 14 us of a 42 us kernel in profiles/r01o)
      float *gdst = dst0 + (size_t)(g * 128 + qd * 32) * kD + chalf * 64;
      const int piece = lane & 15;
      float4 old[16];
      if (!first) {
#pragma unroll
        for (int j = 0; j < 16; ++j) old[j] = *reinterpret_cast<const float4 *>(gdst + (size_t)(2 * j + (lane >> 4)) * kD + piece * 4);
      }
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        float a[16];
        if (my_tiles > 0) {
          tmem_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(g * 128 + chalf * 64 + cc * 16), a);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int x = 0; x < 16; ++x) a[x] = 0.f;
        }
#pragma unroll
        for (int x4 = 0; x4 < 4; ++x4)
          *reinterpret_cast<float4 *>(stg + lane * kLd + cc * 16 + x4 * 4) = make_float4(a[x4 * 4], a[x4 * 4 + 1], a[x4 * 4 + 2], a[x4 * 4 + 3]);
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int row = 2 * j + (lane >> 4);
        float4 v = *reinterpret_cast<const float4 *>(stg + row * kLd + piece * 4);
        if (!first) f4_add(v, old[j]);
        *reinterpret_cast<float4 *>(gdst + (size_t)row * kD + piece * 4) = v;
      }
      __syncwarp();
    }
    if (warp == 2 && lane == 0) trace_stamp(tron, 0, 10);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dW'[384,128] += sum_cta partial[0][cta] ; dWhh += sum_cta partial[1][cta]     (once per backward pass)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float *__restrict__ partial, int ctas, float *__restrict__ dw_fold,
                                                           float *__restrict__ dw_hh) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;        // float4 index inside one [384 x 128] matrix
  const int role = blockIdx.y;
  if (i >= (int)(kWgPartialFloats / 4)) return;
  const float4 *src = reinterpret_cast<const float4 *>(partial + (size_t)role * ctas * kWgPartialFloats) + i;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = 0; c < ctas; ++c) f4_add(s, src[(size_t)c * (kWgPartialFloats / 4)]);
  float4 *dst = reinterpret_cast<float4 *>(role == 0 ? dw_fold : dw_hh) + i;
  float4 d = *dst;
  f4_add(d, s);
  *dst = d;
}

}  // namespace tc2b

// workspace = [dgrad per-slice transposed weight images (384 KB)][q images x4][h image][wgrad partial sums: 2 x 74 x 384 x 128 fp32]
constexpr int kWgCtas = kNumSMs / 2;
static size_t wg_partial_bytes() { return (size_t)2 * kWgCtas * tc2b::kWgPartialFloats * sizeof(float); }
int gru_tc2b_trace_enable(int on) {
  DDFA_CUDA(cudaMemcpyToSymbol(tcc::g_trace_on, &on, sizeof(int)));
  return DDFA_OK;
}
int gru_tc2b_trace_read(void *host, size_t bytes) {
  if (bytes > tcc::kTraceWords * sizeof(long long)) bytes = tcc::kTraceWords * sizeof(long long);
  DDFA_CUDA(cudaMemcpyFromSymbol(host, tcc::g_trace, bytes));
  return DDFA_OK;
}

// workspace: [dgrad3 packed weights 384 KB][h image][dh' * z plane (image-sized)][s image (fp32-s entry only)]
//            [wgrad partial sums][q images x 4] x slots   (one slot, or one per time step when the weight-gradient GEMM of a
//            whole backward pass is batched into one launch)
static constexpr size_t kPackedTotal = tc2b::kD3PackedBytes;
static size_t bwd_fixed_bytes(int32_t N) { return kPackedTotal + 3 * tcc::image_bytes(N) + wg_partial_bytes(); }
void *gru_tc2_bwd_s_image_scratch(void *workspace, int32_t N) { return static_cast<uint8_t *>(workspace) + kPackedTotal + 2 * tcc::image_bytes(N); }
size_t gru_tc2_bwd_workspace_bytes(int32_t N, int32_t slots) {
  return bwd_fixed_bytes(N) + (size_t)(slots < 1 ? 1 : slots) * 4 * tcc::image_bytes(N);
}


int gru_tc2_prepare_bwd(const float *w_fold, const float *w_hh, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (workspace == nullptr || workspace_bytes < kPackedTotal) {
    set_error("tcgen05 engine (bwd): workspace too small");
    return DDFA_ERR_WORKSPACE;
  }
  const int total = 2 * tc2b::kD3Chunks * 128;
  tc2b::dgrad3_pack_kernel<<<(total + 127) / 128, 128, 0, stream>>>(w_fold, w_hh, static_cast<uint32_t *>(workspace));
  DDFA_CHECK_LAUNCH("tc2b::dgrad3_pack_kernel");
  chain_break();
  return DDFA_OK;
}

// dW' += sum of the per-CTA partial sums, dWhh likewise (closes a deferred weight-gradient accumulation)
int gru_tc2_bwd_finish(int32_t N, float *dw_fold, float *dw_hh, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (workspace == nullptr || workspace_bytes < gru_tc2_bwd_workspace_bytes(N, 1)) {
    set_error("tcgen05 engine (bwd finish): workspace too small");
    return DDFA_ERR_WORKSPACE;
  }
  float *partial = reinterpret_cast<float *>(static_cast<uint8_t *>(workspace) + kPackedTotal + 3 * tcc::image_bytes(N));
  const int n4 = (int)(tc2b::kWgPartialFloats / 4);
  tc2b::wgrad_reduce_kernel<<<dim3((n4 + 255) / 256, 2), 256, 0, stream>>>(partial, kWgCtas, dw_fold, dw_hh);
  DDFA_CHECK_LAUNCH("tc2b::wgrad_reduce_kernel");
  return DDFA_OK;
}

// wgrad_mode: 0 = immediate (dW += this step's contribution before returning), 1 = first step of a deferred accumulation
// (partials overwritten), 2 = further deferred step (partials accumulated); deferred passes end with gru_tc2_bwd_finish.
// wgrad_mode >= 16: keep this step's q images in workspace slot (wgrad_mode - 16) and run no weight-gradient GEMM now —
// gru_tc2_bwd_wgrad_batched does it for all kept steps in one launch.
// h_img_in: the activation image of h (kept from the forward pass) or NULL (then it is rebuilt inside the workspace).
// ds_in / indptr_t / indices_t: NULL, or the incoming gradient is dh_out + A^T ds_in (A^T as a CSR over the transposed graph)
int gru_tc2_step_bwd(const float *dh_out, const float *ds_in, const int32_t *indptr_t, const int32_t *indices_t, const float *h,
                     const void *h_img_in, const void *s_img, const float *gates, const void *gates_packed, const int32_t *indptr, int32_t N, float *ds, float *dh, float *dw_fold, float *db_fold, float *db_ih,
                     float *dw_hh, float *db_hh, void *workspace, size_t workspace_bytes, int wgrad_mode, cudaStream_t stream) {
  if (h == nullptr && h_img_in == nullptr) {
    set_error("tcgen05 engine (bwd): neither the fp32 h nor its activation image given");
    return DDFA_ERR_INVALID_ARG;
  }
  const int q_slot = wgrad_mode >= 16 ? wgrad_mode - 16 : 0;
  if (workspace == nullptr || workspace_bytes < gru_tc2_bwd_workspace_bytes(N, q_slot + 1)) {
    set_error("tcgen05 engine (bwd): workspace too small (%zu < %zu)", workspace_bytes, gru_tc2_bwd_workspace_bytes(N, q_slot + 1));
    return DDFA_ERR_WORKSPACE;
  }
  uint8_t *packed = static_cast<uint8_t *>(workspace);
  const size_t img = tcc::image_bytes(N);
  uint8_t *h_img_ws = packed + kPackedTotal;
  float *dhz = reinterpret_cast<float *>(h_img_ws + img);
  float *partial = reinterpret_cast<float *>(h_img_ws + 3 * img);
  uint8_t *q_img = packed + bwd_fixed_bytes(N) + (size_t)q_slot * 4 * img;
  const uint8_t *h_img = h_img_in ? static_cast<const uint8_t *>(h_img_in) : h_img_ws;
  const int64_t rows = ((int64_t)N + tcc::kTileM - 1) / tcc::kTileM * tcc::kTileM;
  unsigned gb_grid = 1;
  {
    const int64_t want = (rows + tc2b::kGbWarps - 1) / tc2b::kGbWarps;
    gb_grid = (unsigned)(want < 2 * kNumSMs ? want : 2 * kNumSMs);
  }
  if (gates_packed && gate_bwd_tma()) {
    // TMA-staged form (packed saved state only): one CTA per SM, three 64 KB stages
    const int blocks32 = (int)(rows / tc2b::kGtRows);
    const int grid = blocks32 < kNumSMs ? blocks32 : kNumSMs;
    const bool csrp = gate_bwd_tma() >= 2;
#define DDFA_GT_LAUNCH(HF32, CSRP)                                                                                                             \
  do {                                                                                                                                         \
    DDFA_CUDA(cudaFuncSetAttribute(tc2b::gate_bwd_tma_kernel<HF32, CSRP>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2b::kGtSmem));        \
    DDFA_CUDA(launch_chain(4, tc2b::gate_bwd_tma_kernel<HF32, CSRP>, dim3(grid), dim3(tc2b::kGtThreads), tc2b::kGtSmem, stream, dh_out, h,     \
                           static_cast<const uint8_t *>(h_img_in), static_cast<const uint2 *>(gates_packed), indptr, ds_in,                    \
                           ds_in ? indptr_t : nullptr, indices_t, N, q_img, img, dhz, db_fold, db_ih, db_hh, l2_hints()));                    \
  } while (0)
    if (h) { if (csrp) DDFA_GT_LAUNCH(true, true); else DDFA_GT_LAUNCH(true, false); }
    else   { if (csrp) DDFA_GT_LAUNCH(false, true); else DDFA_GT_LAUNCH(false, false); }
#undef DDFA_GT_LAUNCH
  } else
  DDFA_CUDA(launch_chain(4, tc2b::gate_bwd_image_kernel, dim3(gb_grid), dim3(32 * tc2b::kGbWarps), 0, stream, dh_out, h,
                         static_cast<const uint8_t *>(h_img_in), gates, static_cast<const uint4 *>(gates_packed), indptr, ds_in,
                         ds_in ? indptr_t : nullptr, indices_t, N, q_img, img, h_img_in ? nullptr : h_img_ws, dhz, db_fold, db_ih, db_hh,
                         l2_hints()));
  DDFA_CHECK_LAUNCH("tc2b::gate_bwd_image_kernel");
  DDFA_CUDA(cudaFuncSetAttribute(tc2b::wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2b::kWgSmemAlloc));
  const int tiles = (N + tcc::kTileM - 1) / tcc::kTileM;
  {
    DDFA_CUDA(cudaFuncSetAttribute(tc2b::dgrad3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2b::kD3SmemAlloc));
    int groups = kNumSMs / 2;
    if (groups > tiles) groups = tiles;
    DDFA_CUDA(launch_chain(8, tc2b::dgrad3_kernel, dim3(groups * 2), dim3(tc2b::kThreads), tc2b::kD3SmemAlloc, stream, q_img, img, dhz,
                           reinterpret_cast<const uint32_t *>(packed), N, ds, dh, l2_hints()));
    DDFA_CHECK_LAUNCH("tc2b::dgrad3_kernel");
  }
  if (wgrad_mode >= 16) return DDFA_OK;       // q images kept; the batched weight-gradient launch follows the last step
  // every one of the 74 x 2 CTAs writes its partial slot (zeros if it owns no tile), so the reduction can sum all of them
  tc2b::WgBatch one = {};
  one.s_img[0] = static_cast<const uint8_t *>(s_img);
  one.h_img[0] = h_img;
  one.steps = 1;
  tc2b::wgrad_kernel<<<dim3(kWgCtas, 2), tc2b::kThreads, tc2b::kWgSmemAlloc, stream>>>(q_img, img, 0, one, N, partial, wgrad_mode == 2 ? 0 : 1, 0);
  DDFA_CHECK_LAUNCH("tc2b::wgrad_kernel");
  if (wgrad_mode == 0) return gru_tc2_bwd_finish(N, dw_fold, dw_hh, workspace, workspace_bytes, stream);
  return DDFA_OK;
}

// dW' += sum_t [q_r q_z q_n]_t^T s_t, dWhh += sum_t [q_r q_z q_nr]_t^T h_t over the `steps` slots kept by gru_tc2_step_bwd
// (wgrad_mode = 16 + slot): ONE weight-gradient launch with K = steps x nodes, then the reduction of the per-CTA partials.
int gru_tc2_bwd_wgrad_batched(const void *const *s_imgs, const void *const *h_imgs, int32_t steps, int32_t N, float *dw_fold,
                              float *dw_hh, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (steps < 1 || steps > tc2b::kWgMaxSteps) {
    set_error("tcgen05 engine (batched wgrad): 1 <= steps <= %d required (got %d)", tc2b::kWgMaxSteps, steps);
    return DDFA_ERR_INVALID_ARG;
  }
  if (workspace == nullptr || workspace_bytes < gru_tc2_bwd_workspace_bytes(N, steps)) {
    set_error("tcgen05 engine (batched wgrad): workspace too small (%zu < %zu)", workspace_bytes, gru_tc2_bwd_workspace_bytes(N, steps));
    return DDFA_ERR_WORKSPACE;
  }
  uint8_t *packed = static_cast<uint8_t *>(workspace);
  const size_t img = tcc::image_bytes(N);
  float *partial = reinterpret_cast<float *>(packed + kPackedTotal + 3 * img);
  const uint8_t *q_img = packed + bwd_fixed_bytes(N);
  tc2b::WgBatch b = {};
  for (int t = 0; t < steps; ++t) {
    if (!s_imgs[t] || !h_imgs[t]) {
      set_error("tcgen05 engine (batched wgrad): NULL image pointer for step %d", t);
      return DDFA_ERR_INVALID_ARG;
    }
    b.s_img[t] = static_cast<const uint8_t *>(s_imgs[t]);
    b.h_img[t] = static_cast<const uint8_t *>(h_imgs[t]);
  }
  b.steps = steps;
  DDFA_CUDA(cudaFuncSetAttribute(tc2b::wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2b::kWgSmemAlloc));
  tc2b::wgrad_kernel<<<dim3(kWgCtas, 2), tc2b::kThreads, tc2b::kWgSmemAlloc, stream>>>(q_img, img, 4 * img, b, N, partial, 1, l2_hints());
  DDFA_CHECK_LAUNCH("tc2b::wgrad_kernel");
  return gru_tc2_bwd_finish(N, dw_fold, dw_hh, workspace, workspace_bytes, stream);
}

}  // namespace ddfa
