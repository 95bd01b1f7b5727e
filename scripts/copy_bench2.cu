// Microbenchmark 2: what bounds the operand feed of the weight-stationary GEMM kernels?
// The in-kernel timeline (profiles/r01l_trace_*.log) shows 32 KB cp.async.bulk copies landing ~0.9 us apart per SM with two
// in flight — half the rate copy_bench.cu measured for a hot private window.  This sweep separates the candidates:
//   * copy size (32 / 64 / 96 KB per copy) and stage count,
//   * one issuing thread vs two issuing threads (different warps),
//   * 1-D bulk copy vs tensor-map (2-D / 3-D tiled) TMA of the same bytes,
//   * access pattern: private hot window | groups of 4 CTAs streaming the same tiles of a large buffer (the kernels' pattern).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o deepdfa_b200/lib/copy_bench2 scripts/copy_bench2.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW_L:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra W_D;\nbra W_L;\nW_D:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap *map, int x, int y, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(x), "r"(y), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_3d(uint32_t dst, const CUtensorMap *map, int x, int y, int z, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
               "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar)
               : "memory");
}

// mode: 0 = 1-D bulk, 1 = 2-D tensor map (boxes of 32 KB: 128 B x 256 rows; chunk / 32 KB boxes per stage),
//       2 = 3-D tensor map (ONE box per chunk: 128 B x 256 rows x chunk / 32 KB)
// pattern: 0 = private window (hot), 1 = groups of `share` CTAs stream chunk ids group + c * ngroups through `total` bytes
// issuers: 1 or 2 threads (warp 0 and warp 2), stage s issued by thread s % issuers
__global__ void __launch_bounds__(128) ring_kernel(const uint8_t *src, const __grid_constant__ CUtensorMap map2, const __grid_constant__ CUtensorMap map3,
                                                   size_t total, size_t window, int chunk, int stages, int mode, int pattern, int share,
                                                   int issuers, int nchunks, unsigned long long *sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)stages * chunk);
  const uint32_t bar0 = smem_u32(bars);
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(bar0 + 8 * i, 1); mbar_init(bar0 + 8 * (stages + i), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int group = blockIdx.x / share, ngroups = gridDim.x / share;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_issuer = lane == 0 && (warp == 0 || (warp == 2 && issuers == 2));
  const int my_issue = warp == 0 ? 0 : 1;
  if (is_issuer) {
    for (int c = 0; c < nchunks; ++c) {
      const int st = c % stages, use = c / stages;
      if (st % issuers != my_issue) continue;
      if (use > 0) mbar_wait(bar0 + 8 * (stages + st), (use - 1) & 1);
      mbar_expect(bar0 + 8 * st, chunk);
      size_t off;
      if (pattern == 0) off = (size_t)blockIdx.x * window + ((size_t)c * chunk) % window;
      else off = (((size_t)group + (size_t)c * ngroups) * chunk) % total;
      const uint32_t dst = smem_u32(smem) + st * chunk;
      if (mode == 0) bulk_g2s(dst, src + off, chunk, bar0 + 8 * st);
      else if (mode == 1) {
        for (int b = 0; b < chunk / 32768; ++b) tma_2d(dst + b * 32768, &map2, 0, (int)(off / 128) + b * 256, bar0 + 8 * st);
      } else {
        tma_3d(dst, &map3, 0, 0, (int)(off / 32768), bar0 + 8 * st);
      }
    }
  } else if (threadIdx.x == 32) {
    unsigned long long acc = 0;
    for (int c = 0; c < nchunks; ++c) {
      const int st = c % stages, use = c / stages;
      mbar_wait(bar0 + 8 * st, use & 1);
      acc += smem[st * chunk + (c & 1023)];
      mbar_arrive(bar0 + 8 * (stages + st));
    }
    sink[blockIdx.x] = acc;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const size_t total = (size_t)96 << 20;   // 96 MB buffer (fits the 126 MB L2 once warm)
  const size_t window = 512 * 1024;
  uint8_t *src; unsigned long long *sink;
  CK(cudaMalloc(&src, total));
  CK(cudaMemset(src, 1, total));
  CK(cudaMalloc(&sink, sizeof(unsigned long long) * 148));
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  CUtensorMap map2, map3[4];
  {
    cuuint64_t dims[2] = {128, total / 128};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {128, 256};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&map2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode 2d failed %d\n", (int)r); return 1; }
  }
  for (int d = 1; d <= 3; ++d) {   // box depth d: d * 32 KB per op
    cuuint64_t dims[3] = {128, 256, total / 32768};
    cuuint64_t strides[2] = {128, 32768};
    cuuint32_t box[3] = {128, 256, (cuuint32_t)d};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(&map3[d], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode 3d depth %d failed %d\n", d, (int)r); return 1; }
  }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
  const int grid = 148;
  const char *mode_name[3] = {"bulk-1D", "tensor-2D(32K boxes)", "tensor-3D(one box)"};
  const char *pat_name[2] = {"private hot window", "4 CTAs share, stream 96 MB"};
  for (int pattern : {0, 1})
    for (int mode : {0, 1, 2})
      for (int chunk : {32768, 65536, 98304})
        for (int stages : {1, 2, 3, 4, 6}) {
          if ((size_t)chunk * stages > 196 * 1024) continue;
          for (int issuers : {1, 2}) {
            if (issuers == 2 && stages < 2) continue;
            const size_t smem = (size_t)chunk * stages + 16 * stages + 64 + 1024;
            const int nchunks = (int)((size_t)(8 << 20) / chunk);   // 8 MB per CTA
            float ms = 0;
            for (int rep_i = 0; rep_i < 2; ++rep_i) {
              CK(cudaEventRecord(e0));
              ring_kernel<<<grid, 128, smem>>>(src, map2, map3[chunk / 32768], total, window, chunk, stages, mode, pattern, pattern ? 4 : 1, issuers,
                                               nchunks, sink);
              CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
              CK(cudaGetLastError());
              CK(cudaEventElapsedTime(&ms, e0, e1));
            }
            const double bytes = (double)nchunks * chunk;
            printf("%-28s | %-22s | chunk %6d x %d stages, %d issuer(s): %8.1f us  %6.3f us/copy  %6.1f GB/s per SM  %6.2f TB/s aggregate\n", pat_name[pattern],
                   mode_name[mode], chunk, stages, issuers, ms * 1e3, ms * 1e3 / nchunks, bytes / (ms * 1e-3) / 1e9, bytes * grid / (ms * 1e-3) / 1e12);
          }
        }
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
