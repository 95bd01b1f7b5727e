// tcgen05 engine, forward: activation-image utilities and the entry points of the forward GRU step.
// The GEMM kernel itself is in gru_tc_fwd3.cu (weights in tensor memory, activations streamed as the B operand).
//
// History, with the measurements that drove it (profiles/):
//   v1 (r01c)  every CTA re-streamed all 384 KB of split weights per tile and converted its operands itself; load / MMA /
//              epilogue serialised — tensor pipe 7-10 % active.
//   v2 (r01e-r01l) weight-stationary: a CTA kept a 96 KB weight slice in shared memory, operands arrived as "activation
//              images" (tc_common.cuh) over cp.async.bulk, four accumulator tiles in flight.  The in-kernel timeline
//              (r01l_trace_fwd.log) showed two limits: with only 2 x 32 KB of shared memory left for operands the feed ran at
//              ~35 GB/s per SM, and every tcgen05.mma issued from an `if (lane == 0)` branch was wrapped by ptxas in a
//              per-instruction election loop (~140 clk per MMA).  A cluster-of-4 multicast feed was tried and measured
//              slower (r01k_tcdebug_cluster_ab.log: 56.8 vs 43.3 us; at cluster size 4 the L2 already de-duplicates).
//   v3 (r01p-)  gru_tc_fwd3.cu.
#include "tc_common.cuh"

namespace ddfa {
namespace tc2 {
using namespace tcc;

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

size_t gru_tc2_workspace_bytes() { return gru_tc3_packed_bytes(); }

int gru_tc2_prepare(const float *w_fold, const float *b_fold, const float *b_ih, const float *w_hh, const float *b_hh,
                    void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (workspace == nullptr || workspace_bytes < gru_tc2_workspace_bytes()) {
    set_error("tcgen05 engine: workspace too small (%zu < %zu)", workspace_bytes, gru_tc2_workspace_bytes());
    return DDFA_ERR_WORKSPACE;
  }
  return gru_tc3_prepare(w_fold, b_fold, b_ih, w_hh, b_hh, workspace, stream);
}

int gru_tc2_step_fwd(const void *s_img, const void *h_img, const float *h, const int32_t *indptr, int32_t N, float *h_out,
                     void *h_out_img, float *save_gates, const void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (workspace == nullptr || workspace_bytes < gru_tc2_workspace_bytes()) {
    set_error("tcgen05 engine: workspace too small (%zu < %zu)", workspace_bytes, gru_tc2_workspace_bytes());
    return DDFA_ERR_WORKSPACE;
  }
  return gru_tc3_step_fwd(s_img, h_img, h, indptr, N, h_out, h_out_img, save_gates, workspace, stream);
}

}  // namespace ddfa
