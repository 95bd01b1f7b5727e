// K3 — CSR edge gather-sum:  out[v,:] = (acc ? out[v,:] : 0) + sum_{e in [indptr[v], indptr[v+1])} h[indices[e], :]
//
// Replaces DGL's update_all(fn.copy_u('h','m'), fn.sum('m','a')) (SpMM) inside GatedGraphConv
// (reference call site DDFA/code_gnn/models/flow_gnn/ggnn.py:95).  Run on the CSR of the
// transposed graph it is that op's backward.
//
// Bound: HBM bandwidth (zero FLOPs).  Algorithmic bytes per launch:
//     E*D*4 (source rows) + N*D*4 (write) + E*4 (indices) + (N+1)*4 (indptr)
//
// Mapping: a group of G = min(32, D/4) lanes owns RW consecutive destination rows; one lane
// holds one 16-byte column chunk, so a D=128 fp32 row (512 B) is exactly one warp-wide
// ld.global.nc.v4.  Per group: one coalesced load of the RW+1 row pointers, one coalesced load
// of the (<=32 per pass) neighbour ids, then the neighbour-row loads are issued in batches of
// UNROLL independent 128-bit loads (memory-level parallelism) and folded into the per-row
// accumulators by a uniform segmented reduction (row boundaries broadcast with shuffles).
#include "common.cuh"

namespace ddfa {

template <int G, int CH, int RW, int UNROLL>
__global__ void __launch_bounds__(256) gather_sum_kernel(const int32_t *__restrict__ indptr,
                                                         const int32_t *__restrict__ indices,
                                                         const float *__restrict__ h, int32_t N, int32_t D,
                                                         float *__restrict__ out, int accumulate) {
  constexpr int GROUPS_PER_WARP = 32 / G;
  const int lane = threadIdx.x & 31;
  const int gl = lane % G;                 // lane inside the group
  const int gbase = lane - gl;             // first lane of the group inside the warp
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gbase);
  const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t group = warp_global * GROUPS_PER_WARP + (lane / G);
  const int64_t v0 = group * RW;
  if (v0 >= N) return;
  const int nrows = (int)min((int64_t)RW, (int64_t)N - v0);

  // row pointers of this chunk: lane i holds indptr[v0+i], i <= nrows
  int32_t myptr = 0;
  if (gl <= nrows) myptr = __ldg(indptr + v0 + gl);
  const int32_t beg0 = __shfl_sync(gmask, myptr, gbase);
  int32_t rend[RW];  // row end offsets relative to beg0
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    int32_t e = __shfl_sync(gmask, myptr, gbase + min(r + 1, nrows));
    rend[r] = e - beg0;
  }
  const int32_t total = rend[RW - 1];

  float4 acc[RW][CH];
#pragma unroll
  for (int r = 0; r < RW; ++r)
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[r][c] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int32_t pbase = 0; pbase < total; pbase += G) {
    // coalesced load of up to G neighbour ids of this chunk
    int32_t myidx = 0;
    if (pbase + gl < total) myidx = __ldg(indices + beg0 + pbase + gl);
    const int32_t cnt = min((int32_t)G, total - pbase);
    for (int32_t b = 0; b < cnt; b += UNROLL) {
      float4 v[UNROLL][CH];
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) {
        const int32_t u = __shfl_sync(gmask, myidx, gbase + min(b + j, cnt - 1));
        if (b + j < cnt) {
          const float *row = h + (int64_t)u * D;
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const int col = (gl + c * G) * 4;
            v[j][c] = (col < D) ? ldg_nc_f4(row + col) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        } else {
#pragma unroll
          for (int c = 0; c < CH; ++c) v[j][c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) {
        const int32_t pos = pbase + b + j;  // position inside the chunk's edge list
        // uniform (per group) segmented accumulate: edge `pos` belongs to the first row with rend > pos
#pragma unroll
        for (int r = 0; r < RW; ++r) {
          const bool mine = (pos < rend[r]) && (r == 0 || pos >= rend[r - 1]);
          if (mine) {
#pragma unroll
            for (int c = 0; c < CH; ++c) f4_add(acc[r][c], v[j][c]);
          }
        }
      }
    }
  }

#pragma unroll
  for (int r = 0; r < RW; ++r) {
    if (r < nrows) {
      float *orow = out + (v0 + r) * D;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int col = (gl + c * G) * 4;
        if (col < D) {
          float4 a = acc[r][c];
          if (accumulate) f4_add(a, *reinterpret_cast<const float4 *>(orow + col));
          *reinterpret_cast<float4 *>(orow + col) = a;
        }
      }
    }
  }
}

template <int G, int CH, int RW, int UNROLL>
static int launch_gather(const int32_t *indptr, const int32_t *indices, const float *h, int32_t N, int32_t D,
                         float *out, int accumulate, cudaStream_t stream) {
  constexpr int GROUPS_PER_WARP = 32 / G;
  const int64_t groups = ((int64_t)N + RW - 1) / RW;
  const int64_t warps = (groups + GROUPS_PER_WARP - 1) / GROUPS_PER_WARP;
  const int threads = 256;
  const int64_t blocks = (warps * 32 + threads - 1) / threads;
  gather_sum_kernel<G, CH, RW, UNROLL><<<(unsigned)blocks, threads, 0, stream>>>(indptr, indices, h, N, D, out, accumulate);
  DDFA_CHECK_LAUNCH("gather_sum_kernel");
  return DDFA_OK;
}

}  // namespace ddfa

extern "C" int ddfa_gather_sum(const int32_t *indptr, const int32_t *indices, const float *h, int32_t N,
                               int32_t D, float *out, int accumulate, void *stream_) {
  using namespace ddfa;
  DDFA_REQUIRE(N >= 0 && D > 0 && D % 4 == 0 && D <= 1024, "ddfa_gather_sum: unsupported shape N=%d D=%d (need D%%4==0, D<=1024)", N, D);
  if (N == 0) return DDFA_OK;
  DDFA_REQUIRE(indptr && indices && h && out, "ddfa_gather_sum: NULL pointer");
  DDFA_REQUIRE(aligned16(h) && aligned16(out), "ddfa_gather_sum: h/out must be 16-byte aligned");
  DDFA_REQUIRE(h != out, "ddfa_gather_sum: in-place gather is not supported");
  cudaStream_t stream = as_stream(stream_);
  const int chunks = D / 4;  // 16-byte chunks per row
  if (chunks <= 8) return launch_gather<8, 1, 4, 8>(indptr, indices, h, N, D, out, accumulate, stream);
  if (chunks <= 16) return launch_gather<16, 1, 4, 8>(indptr, indices, h, N, D, out, accumulate, stream);
  if (chunks <= 32) return launch_gather<32, 1, 4, 8>(indptr, indices, h, N, D, out, accumulate, stream);
  if (chunks <= 64) return launch_gather<32, 2, 2, 8>(indptr, indices, h, N, D, out, accumulate, stream);
  if (chunks <= 128) return launch_gather<32, 4, 2, 4>(indptr, indices, h, N, D, out, accumulate, stream);
  return launch_gather<32, 8, 1, 4>(indptr, indices, h, N, D, out, accumulate, stream);
}
