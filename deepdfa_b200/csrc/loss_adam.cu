// K8 — graph labels + BCE-with-logits loss (+ d loss / d logits), and K10 — fused Adam.
//
// K8 replaces BaseModule.get_label (base_module.py:83-95: dgl.unbatch + a Python loop taking
// max(_VULN) per graph) and torch.nn.BCEWithLogitsLoss(pos_weight) (base_module.py:72-74,183).
// K10 replaces torch.optim.Adam(lr=1e-3, weight_decay=1e-2) — coupled L2, not AdamW
// (DDFA/configs/config_default.yaml:43-47) — over one flat parameter buffer.
#include <math.h>

#include "common.cuh"

namespace ddfa {

// warp per graph: segment max of vuln, then the loss term of that graph
__global__ void __launch_bounds__(256) graph_label_bce_kernel(const float *__restrict__ logits, const int32_t *__restrict__ vuln,
                                                              const int32_t *__restrict__ graph_ptr, int32_t B, int32_t B_valid,
                                                              float pos_weight, float loss_scale, float grad_scale, float *__restrict__ labels,
                                                              float *__restrict__ loss_out, float *__restrict__ dlogits) {
  __shared__ float s_loss[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * 8 + warp;
  float term = 0.f;
  if (b < B) {
    const int32_t n0 = graph_ptr[b], n1 = graph_ptr[b + 1];
    int32_t mx = INT32_MIN;
    for (int32_t n = n0 + lane; n < n1; n += 32) mx = max(mx, vuln[n]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (n1 <= n0) mx = 0;  // empty graph: no label information
    const float y = (float)mx;
    if (lane == 0) {
      if (labels) labels[b] = y;
      if (dlogits && b >= B_valid) dlogits[b] = 0.f;      // padding graphs (shape bucketing): no loss term, no gradient
      if (logits && b < B_valid) {
        const float x = logits[b];
        // torch: (1-y)*x + (1+(pw-1)*y) * (log1p(exp(-|x|)) + max(-x,0))
        const float lw = 1.f + (pos_weight - 1.f) * y;
        term = (1.f - y) * x + lw * (log1pf(expf(-fabsf(x))) + fmaxf(-x, 0.f));
        if (dlogits) {
          const float sg = 1.f / (1.f + expf(-x));
          // d/dx = (1-y) - lw * (1 - sigmoid(x)) = sigmoid(x)*lw - y*pw ... expanded for clarity:
          dlogits[b] = grad_scale * ((1.f - y) - lw * (1.f - sg));
        }
      }
    }
  }
  if (lane == 0) s_loss[warp] = term;
  __syncthreads();
  if (threadIdx.x == 0 && loss_out) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += s_loss[w];
    atomicAdd(loss_out, loss_scale * s);
  }
}

__global__ void __launch_bounds__(256) adam_flat_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                                        float *__restrict__ v, const int32_t *__restrict__ step_count, int64_t n,
                                                        float lr, float beta1, float beta2, float eps, float wd) {
  __shared__ float s_c[2];
  if (threadIdx.x == 0) {
    const double t = (double)(*step_count + 1);
    const double bc1 = 1.0 - pow((double)beta1, t);
    const double bc2 = 1.0 - pow((double)beta2, t);
    s_c[0] = (float)((double)lr / bc1);   // step_size
    s_c[1] = (float)sqrt(bc2);            // bias_correction2_sqrt
  }
  __syncthreads();
  const float step_size = s_c[0], bc2s = s_c[1];
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = g[i];
  const float pi = p[i];
  gi = fmaf(wd, pi, gi);                               // grad = grad + wd * param  (coupled L2)
  const float mi = fmaf(beta1, m[i], (1.f - beta1) * gi);  // exp_avg.lerp_(grad, 1-beta1)
  const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2s + eps;
  p[i] = pi - step_size * (mi / denom);
}
__global__ void adam_step_inc_kernel(int32_t *step_count) { *step_count += 1; }

int adam_step_inc_launch(int32_t *step_count, cudaStream_t stream) {
  adam_step_inc_kernel<<<1, 1, 0, stream>>>(step_count);
  DDFA_CHECK_LAUNCH("adam_step_inc_kernel");
  return DDFA_OK;
}

}  // namespace ddfa

extern "C" {

int ddfa_graph_label_bce_valid(const float *logits, const int32_t *vuln, const int32_t *graph_ptr, int32_t B, int32_t B_valid,
                               float pos_weight, float loss_scale, float grad_scale, float *labels, float *loss_out, float *dlogits,
                               void *stream_) {
  using namespace ddfa;
  DDFA_REQUIRE(B >= 0 && B_valid >= 0 && B_valid <= B, "ddfa_graph_label_bce: need 0 <= num_valid (%d) <= num_graphs (%d)", B_valid, B);
  if (B == 0) return DDFA_OK;
  DDFA_REQUIRE(vuln && graph_ptr, "ddfa_graph_label_bce: NULL pointer");
  DDFA_REQUIRE(logits || (!loss_out && !dlogits), "ddfa_graph_label_bce: loss requested without logits");
  cudaStream_t stream = as_stream(stream_);
  if (loss_out) DDFA_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), stream));
  graph_label_bce_kernel<<<(B + 7) / 8, 256, 0, stream>>>(logits, vuln, graph_ptr, B, B_valid, pos_weight, loss_scale, grad_scale, labels,
                                                         loss_out, dlogits);
  DDFA_CHECK_LAUNCH("graph_label_bce_kernel");
  return DDFA_OK;
}

int ddfa_graph_label_bce(const float *logits, const int32_t *vuln, const int32_t *graph_ptr, int32_t B, float pos_weight,
                         float loss_scale, float grad_scale, float *labels, float *loss_out, float *dlogits, void *stream_) {
  return ddfa_graph_label_bce_valid(logits, vuln, graph_ptr, B, B, pos_weight, loss_scale, grad_scale, labels, loss_out, dlogits, stream_);
}

int ddfa_adam_flat(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, int32_t *step_count, int64_t numel,
                   float lr, float beta1, float beta2, float eps, float weight_decay, void *stream_) {
  using namespace ddfa;
  DDFA_REQUIRE(numel >= 0, "ddfa_adam_flat: negative numel");
  DDFA_REQUIRE(params && grads && exp_avg && exp_avg_sq && step_count, "ddfa_adam_flat: NULL pointer");
  cudaStream_t stream = as_stream(stream_);
  if (numel > 0) {
    adam_flat_kernel<<<(unsigned)((numel + 255) / 256), 256, 0, stream>>>(params, grads, exp_avg, exp_avg_sq, step_count, numel, lr,
                                                                         beta1, beta2, eps, weight_decay);
    DDFA_CHECK_LAUNCH("adam_flat_kernel");
  }
  return adam_step_inc_launch(step_count, stream);
}

}  // extern "C"
