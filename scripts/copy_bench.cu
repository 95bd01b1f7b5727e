// Microbenchmark: how fast can ONE CTA per SM stream L2-resident data into shared memory on B200?
// Compares the feed mechanisms available to the GEMM kernels:
//   mode 0: cp.async.bulk 1-D (UBLKCP) ring, one issuing thread, chunk bytes / stages swept
//   mode 1: cp.async 16 B (LDGSTS) issued by W warps, ring of stages, mbarrier completion
//   mode 2: ld.global.v4 -> st.shared.v4 by W warps (register staged), __syncthreads per chunk
// Each CTA reads its own 1 MB window (so 148 MB total, > L2? no: L2 is 126 MB -> use 0.5 MB windows = 74 MB, L2-resident
// after the first pass) `iters` times.  Prints GB/s per SM and aggregate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o deepdfa_b200/lib/copy_bench scripts/copy_bench.cu
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

// mode 0: bulk ring.  warp 0 lane 0 produces, warp 1 lane 0 "consumes" (waits full, arrives empty).
__global__ void __launch_bounds__(128) bulk_ring_kernel(const uint8_t *src, size_t window, int chunk, int stages, int pieces, int iters, unsigned long long *sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)stages * chunk);
  const uint32_t bar0 = smem_u32(bars);
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(bar0 + 8 * i, 1); mbar_init(bar0 + 8 * (stages + i), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint8_t *base = src + (size_t)blockIdx.x * window;
  const int nchunks = (int)(window / chunk) * iters;
  if (threadIdx.x == 0) {
    for (int c = 0; c < nchunks; ++c) {
      const int st = c % stages, use = c / stages;
      if (use > 0) mbar_wait(bar0 + 8 * (stages + st), (use - 1) & 1);
      mbar_expect(bar0 + 8 * st, chunk);
      const size_t off = ((size_t)c * chunk) % window;
      const int pb = chunk / pieces;
      for (int p = 0; p < pieces; ++p) bulk_g2s(smem_u32(smem) + st * chunk + p * pb, base + off + (size_t)p * pb, pb, bar0 + 8 * st);
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

// mode 1: LDGSTS ring: `lw` loader warps issue 16-byte cp.async, completion via cp.async.mbarrier.arrive.noinc
__global__ void __launch_bounds__(320) ldgsts_ring_kernel(const uint8_t *src, size_t window, int chunk, int stages, int lw, int iters, unsigned long long *sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)stages * chunk);
  const uint32_t bar0 = smem_u32(bars);
  const int nload = lw * 32;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(bar0 + 8 * i, nload); mbar_init(bar0 + 8 * (stages + i), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint8_t *base = src + (size_t)blockIdx.x * window;
  const int nchunks = (int)(window / chunk) * iters;
  const int warp = threadIdx.x >> 5;
  if (warp < lw) {
    for (int c = 0; c < nchunks; ++c) {
      const int st = c % stages, use = c / stages;
      if (use > 0) mbar_wait(bar0 + 8 * (stages + st), (use - 1) & 1);
      const size_t off = ((size_t)c * chunk) % window;
      for (int o = threadIdx.x * 16; o < chunk; o += nload * 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem) + st * chunk + o), "l"(base + off + o) : "memory");
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar0 + 8 * st) : "memory");
    }
  } else if (threadIdx.x == lw * 32) {
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

// mode 2: plain LDG.128 -> STS.128, 8 loads in flight per thread
__global__ void __launch_bounds__(256) ldg_sts_kernel(const uint8_t *src, size_t window, int iters, unsigned long long *sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint8_t *base = src + (size_t)blockIdx.x * window;
  const int per_pass = 256 * 16 * 8;  // 32 KB per pass
  unsigned long long acc = 0;
  for (int it = 0; it < iters; ++it)
    for (size_t off = 0; off < window; off += per_pass) {
      uint4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = *reinterpret_cast<const uint4 *>(base + off + (size_t)(j * 256 + threadIdx.x) * 16);
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4 *>(smem + (size_t)(j * 256 + threadIdx.x) * 16) = v[j];
      __syncthreads();
      acc += smem[threadIdx.x];
      __syncthreads();
    }
  if (threadIdx.x == 0) sink[blockIdx.x] = acc;
}

int main() {
  const int sms = 148;
  const size_t window = 512 * 1024;
  uint8_t *src; unsigned long long *sink;
  CK(cudaMalloc(&src, window * sms));
  CK(cudaMemset(src, 1, window * sms));
  CK(cudaMalloc(&sink, sizeof(unsigned long long) * sms));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 8;
  const double total_bytes = (double)window * sms * iters;
  auto report = [&](const char *name, float ms) {
    printf("%-58s %8.1f us  %7.1f GB/s per SM  %7.2f TB/s aggregate\n", name, ms * 1e3, total_bytes / sms / (ms * 1e-3) / 1e9, total_bytes / (ms * 1e-3) / 1e12);
  };
  CK(cudaFuncSetAttribute(bulk_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(ldgsts_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(ldg_sts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  const int chunks[] = {4096, 8192, 16384, 32768};
  for (int grid : {148, 37}) {
    printf("---- grid = %d CTAs\n", grid);
    const double tb = (double)window * grid * iters;
    auto rep = [&](const char *name, float ms) {
      printf("%-58s %8.1f us  %7.1f GB/s per SM  %7.2f TB/s aggregate\n", name, ms * 1e3, tb / grid / (ms * 1e-3) / 1e9, tb / (ms * 1e-3) / 1e12);
    };
    for (int chunk : chunks)
      for (int stages : {2, 4, 7}) {
        if ((size_t)chunk * stages > 180 * 1024) continue;
        for (int pieces : {1, 4}) {
          const size_t smem = (size_t)chunk * stages + 16 * stages + 64;
          for (int rep_i = 0; rep_i < 2; ++rep_i) {
            CK(cudaEventRecord(e0));
            bulk_ring_kernel<<<grid, 128, smem>>>(src, window, chunk, stages, pieces, iters, sink);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
          }
          float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
          char name[128]; snprintf(name, sizeof(name), "bulk 1-D: chunk %5d B x %d stages, %d piece(s)", chunk, stages, pieces);
          rep(name, ms);
        }
      }
    for (int lw : {2, 4, 8})
      for (int stages : {4, 7}) {
        const int chunk = 16384;
        const size_t smem = (size_t)chunk * stages + 16 * stages + 64;
        for (int rep_i = 0; rep_i < 2; ++rep_i) {
          CK(cudaEventRecord(e0));
          ldgsts_ring_kernel<<<grid, 320, smem>>>(src, window, chunk, stages, lw, iters, sink);
          CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        }
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        char name[128]; snprintf(name, sizeof(name), "LDGSTS 16 B: chunk 16384 B x %d stages, %d loader warps", stages, lw);
        rep(name, ms);
      }
    for (int rep_i = 0; rep_i < 2; ++rep_i) {
      CK(cudaEventRecord(e0));
      ldg_sts_kernel<<<grid, 256, 32 * 1024>>>(src, window, iters, sink);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    }
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    rep("LDG.128 -> STS.128, 256 threads, 8 loads in flight", ms);
  }
  (void)report; (void)total_bytes;
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
