// K1 — embedding lookup + concat, forward and backward.
// Reference: 4x nn.Embedding(input_dim, hidden_dim) + torch.cat(dim=1) (ggnn.py:47-52,84-89) or a
// single nn.Embedding (ggnn.py:54,91-92).  Tables total 4*1002*32*4 B = 513 KB -> L2 resident;
// the op is bound by the index read (8 B/node/table) and the 4*D B/node output write.
#include <cuda_bf16.h>

#include "common.cuh"

namespace ddfa {

constexpr int kMaxTables = 8;
__device__ __forceinline__ uint32_t bf16x2_pack(float lo_half, float hi_half) {     // cvt.rn.bf16x2.f32: .x -> bits 0-15
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo_half, hi_half);
  return *reinterpret_cast<const uint32_t *>(&v);
}
struct EmbedPtrs {
  const int64_t *idx[kMaxTables];
  const float *table[kMaxTables];
  float *dtable[kMaxTables];
};

// one thread per 16-byte output chunk.  IMG (row width 128 only): the row also goes out as h_0's activation image (bf16 hi / lo,
// K-major SWIZZLE_128B tiles, include/ddfa_b200.h) — what a separate ddfa_act_to_image pass over x produced before.
template <bool IMG>
__global__ void __launch_bounds__(256) embed_concat_fwd_kernel(const EmbedPtrs p, int32_t K, int32_t V, int32_t H,
                                                               int32_t N, float *__restrict__ x, uint8_t *__restrict__ image,
                                                               int32_t *__restrict__ oob) {
  const int hq = H >> 2;           // chunks per table
  const int dq = K * hq;           // chunks per node row
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (int64_t)N * dq) return;
  const int32_t n = (int32_t)(t / dq);
  const int c = (int)(t - (int64_t)n * dq);
  const int k = c / hq, j = (c - k * hq) * 4;
  int64_t i = p.idx[k][n];
  if (i < 0 || i >= V) {
    if (oob && j == 0) atomicAdd(oob, 1);
    i = i < 0 ? 0 : V - 1;
  }
  const float4 v = ldg_nc_f4(p.table[k] + i * H + j);
  *reinterpret_cast<float4 *>(x + (int64_t)n * (K * H) + c * 4) = v;
  if constexpr (IMG) {
    const int col = c * 4, row = n & 127;
    const uint32_t h01 = bf16x2_pack(v.x, v.y), h23 = bf16x2_pack(v.z, v.w);
    const uint32_t l01 = bf16x2_pack(v.x - __uint_as_float(h01 << 16), v.y - __uint_as_float(h01 & 0xffff0000u));
    const uint32_t l23 = bf16x2_pack(v.z - __uint_as_float(h23 << 16), v.w - __uint_as_float(h23 & 0xffff0000u));
    uint8_t *tile = image + (size_t)(n >> 7) * 65536;
    const uint32_t sw = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + (((((col & 63) >> 3) ^ (row & 7)) & 7) << 4) + (col & 7) * 2);
    *reinterpret_cast<uint2 *>(tile + (size_t)((0 * 2 + (col >> 6)) * 16384) + sw) = make_uint2(h01, h23);
    *reinterpret_cast<uint2 *>(tile + (size_t)((1 * 2 + (col >> 6)) * 16384) + sw) = make_uint2(l01, l23);
  }
}

// Backward: dtable[k][idx_k[n], :] += (dx + dx2)[n, kH:(k+1)H]   (dx2 optional).
// ~75 % of Big-Vul nodes carry index 0 ("not a definition", dbize_absdf.py:39) and a few % index 1
// (UNKNOWN), so rows 0 and 1 are privatised: each thread accumulates them in registers over its
// rows, the CTA reduces through shared memory and issues ONE RED per element per CTA; all other
// rows go straight to L2 with RED.ADD.F32.
__device__ __forceinline__ void red_add_f4(float *p, const float4 &v) {      // p 16-byte aligned (H % 4 == 0, tables 16-byte aligned)
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

constexpr int kEmbRows = 256;  // node rows per CTA
__global__ void __launch_bounds__(256) embed_concat_bwd_kernel(const EmbedPtrs p, int32_t K, int32_t V, int32_t H,
                                                               int32_t N, const float *__restrict__ dx,
                                                               const float *__restrict__ dx2) {
  extern __shared__ __align__(16) float red[];  // [ny][2][D]
  const int D = K * H;
  const int c = threadIdx.x;  // chunk inside the row (blockDim.x == D/4)
  const int hq = H >> 2;
  const int k = c / hq, j = (c - k * hq) * 4;
  const int64_t *idx = p.idx[k];
  float *dt = p.dtable[k];
  float4 hot[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
  const int32_t row0 = blockIdx.x * kEmbRows;
  const int32_t row1 = min(N, row0 + kEmbRows);
  // four rows per iteration: their (independent) loads are in flight together — one row at a time was a chain of load latencies
  // (r03a launch list: 96 us for 157 MB at C1)
  constexpr int U = 4;
  for (int32_t nb = row0 + threadIdx.y; nb < row1; nb += U * blockDim.y) {
    float4 g[U];
    int64_t ii[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int32_t n = nb + u * blockDim.y;
      const bool ok = n < row1;
      g[u] = ok ? __ldg(reinterpret_cast<const float4 *>(dx + (int64_t)n * D + c * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (dx2 && ok) f4_add(g[u], __ldg(reinterpret_cast<const float4 *>(dx2 + (int64_t)n * D + c * 4)));
      ii[u] = ok ? idx[n] : -1;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (ii[u] == -1 && nb + u * (int32_t)blockDim.y >= row1) continue;
      int64_t i = ii[u];
      i = i < 0 ? 0 : (i >= V ? V - 1 : i);
      if (i == 0) f4_add(hot[0], g[u]);
      else if (i == 1) f4_add(hot[1], g[u]);
      else {
        red_add_f4(dt + i * H + j, g[u]);      // one 16-byte reduction instead of four (sm_90+: red.global.add.v4.f32)
      }
    }
  }
  *reinterpret_cast<float4 *>(&red[((size_t)threadIdx.y * 2 + 0) * D + c * 4]) = hot[0];
  *reinterpret_cast<float4 *>(&red[((size_t)threadIdx.y * 2 + 1) * D + c * 4]) = hot[1];
  __syncthreads();
  if (threadIdx.y < 2) {
    const int r = threadIdx.y;  // hot row 0 or 1
    if (r < V) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int y = 0; y < blockDim.y; ++y) f4_add(s, *reinterpret_cast<const float4 *>(&red[((size_t)y * 2 + r) * D + c * 4]));
      red_add_f4(dt + (int64_t)r * H + j, s);
    }
  }
}

}  // namespace ddfa

extern "C" {

static int embed_fwd_impl(const char *who, const int64_t *const *idx, const float *const *tables, int32_t K, int32_t V, int32_t H, int32_t N,
                          float *x, void *image, int32_t *oob_count, void *stream_) {
  using namespace ddfa;
  DDFA_REQUIRE(K >= 1 && K <= kMaxTables && V > 0 && H > 0 && H % 4 == 0 && N >= 0,
               "%s: unsupported shape K=%d V=%d H=%d N=%d (H%%4==0, K<=%d)", who, K, V, H, N, kMaxTables);
  DDFA_REQUIRE(image == nullptr || K * H == 128, "%s: the activation image exists for row width 128 only (K*H=%d)", who, K * H);
  if (N == 0) return DDFA_OK;
  DDFA_REQUIRE(idx && tables && x && aligned16(x) && aligned16(image), "%s: NULL or unaligned pointer", who);
  EmbedPtrs p{};
  for (int k = 0; k < K; ++k) {
    DDFA_REQUIRE(idx[k] && tables[k] && aligned16(tables[k]), "%s: table %d pointer NULL or unaligned", who, k);
    p.idx[k] = idx[k];
    p.table[k] = tables[k];
  }
  const int64_t tot = (int64_t)N * K * (H / 4);
  const unsigned grid = (unsigned)((tot + 255) / 256);
  if (image) embed_concat_fwd_kernel<true><<<grid, 256, 0, as_stream(stream_)>>>(p, K, V, H, N, x, static_cast<uint8_t *>(image), oob_count);
  else embed_concat_fwd_kernel<false><<<grid, 256, 0, as_stream(stream_)>>>(p, K, V, H, N, x, nullptr, oob_count);
  DDFA_CHECK_LAUNCH("embed_concat_fwd_kernel");
  return DDFA_OK;
}

int ddfa_embed_concat_fwd(const int64_t *const *idx, const float *const *tables, int32_t K, int32_t V, int32_t H,
                          int32_t N, float *x, int32_t *oob_count, void *stream_) {
  return embed_fwd_impl("ddfa_embed_concat_fwd", idx, tables, K, V, H, N, x, nullptr, oob_count, stream_);
}

int ddfa_embed_concat_fwd_image(const int64_t *const *idx, const float *const *tables, int32_t K, int32_t V, int32_t H,
                                int32_t N, float *x, void *image, int32_t *oob_count, void *stream_) {
  DDFA_REQUIRE(image != nullptr, "ddfa_embed_concat_fwd_image: NULL image");
  return embed_fwd_impl("ddfa_embed_concat_fwd_image", idx, tables, K, V, H, N, x, image, oob_count, stream_);
}

int ddfa_embed_concat_bwd(const int64_t *const *idx, const float *dx, const float *dx2, int32_t K, int32_t V, int32_t H,
                          int32_t N, float *const *dtables, void *stream_) {
  using namespace ddfa;
  DDFA_REQUIRE(K >= 1 && K <= kMaxTables && V > 0 && H > 0 && H % 4 == 0 && N >= 0 && K * H <= 512,
               "ddfa_embed_concat_bwd: unsupported shape K=%d V=%d H=%d N=%d", K, V, H, N);
  if (N == 0) return DDFA_OK;
  DDFA_REQUIRE(idx && dx && dtables && aligned16(dx) && aligned16(dx2), "ddfa_embed_concat_bwd: NULL or unaligned pointer");
  for (int k = 0; k < K; ++k) DDFA_REQUIRE(aligned16(dtables[k]), "ddfa_embed_concat_bwd: gradient table %d must be 16-byte aligned", k);
  EmbedPtrs p{};
  for (int k = 0; k < K; ++k) {
    DDFA_REQUIRE(idx[k] && dtables[k], "ddfa_embed_concat_bwd: table %d pointer NULL", k);
    p.idx[k] = idx[k];
    p.dtable[k] = dtables[k];
  }
  const int D = K * H;
  dim3 block(D / 4, (256 / (D / 4)) > 2 ? 256 / (D / 4) : 2);
  const size_t smem = sizeof(float) * block.y * 2 * D;
  embed_concat_bwd_kernel<<<(N + kEmbRows - 1) / kEmbRows, block, smem, as_stream(stream_)>>>(p, K, V, H, N, dx, dx2);
  DDFA_CHECK_LAUNCH("embed_concat_bwd_kernel");
  return DDFA_OK;
}

}  // extern "C"
