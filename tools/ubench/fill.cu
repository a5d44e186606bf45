// Micro-benchmark: bytes per clock an SM receives from L2 through bulk-async (TMA) copies into a shared-memory ring, when every SM
// streams the SAME L2-resident region (the weight stream of a convolution), with and without cluster multicast.
//   ./fill <cluster 1|2|4> <slots> <slot KB> <region KB> <iters> [distinct]
// cluster > 1: every slot fill is split into `cluster` parts; CTA r loads part r and multicasts it to all CTAs of the cluster.
// Output: GB/s delivered per SM and chip-wide, bytes/clk/SM.  Build: nvcc -arch=sm_100a -O3 -o fill fill.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ void arrive_remote(uint32_t a) { asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int CL>
__global__ void __launch_bounds__(64, 1) k_fill(const uint8_t* src, size_t region, size_t per_cta_stride, int slots, uint32_t slot_bytes, int iters,
                                               unsigned long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);          // [16]
  uint64_t* empty = full + 16;                                  // [16]
  uint8_t* ring = smem + 1024;
  const uint32_t rank = CL > 1 ? cluster_rank() : 0u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], CL); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (CL > 1) cluster_sync(); else __syncthreads();
  const uint8_t* base = src + (size_t)(blockIdx.x / CL) * per_cta_stride;     // per_cta_stride = 0: every cluster streams the same region
  const uint32_t part = slot_bytes / CL;
  const long long t0 = clock64();
  if (threadIdx.x == 0) {                                       // producer
    int st = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&empty[st], ph ^ 1);
      mbar_expect_tx(&full[st], slot_bytes);
      const uint8_t* g = base + ((size_t)it * slot_bytes) % region + (size_t)rank * part;
      uint8_t* d = ring + (size_t)st * slot_bytes + (size_t)rank * part;
      if (CL == 1) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(d)), "l"(g), "r"(part),
                     "r"(smem_u32(&full[st])) : "memory");
      } else {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(d)),
                     "l"(g), "r"(part), "r"(smem_u32(&full[st])), "h"((uint16_t)((1u << CL) - 1)) : "memory");
      }
      if (++st == slots) { st = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {                               // consumer: frees the slot in every CTA of the cluster at once
    int st = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full[st], ph);
      if (CL == 1) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[st])) : "memory");
      else for (uint32_t r = 0; r < (uint32_t)CL; ++r) arrive_remote(mapa(smem_u32(&empty[st]), r));
      if (++st == slots) { st = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = (unsigned long long)(clock64() - t0);
  if (CL > 1) cluster_sync();
}

int main(int argc, char** argv) {
  const int cl = argc > 1 ? atoi(argv[1]) : 1, slots = argc > 2 ? atoi(argv[2]) : 5;
  const uint32_t slot_bytes = (argc > 3 ? atoi(argv[3]) : 16) * 1024u;
  const size_t region = (size_t)(argc > 4 ? atoi(argv[4]) : 1024) * 1024;
  const int iters = argc > 5 ? atoi(argv[5]) : 20000;
  const bool distinct = argc > 6;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms / cl * cl;
  const size_t stride = distinct ? region : 0, total = distinct ? region * (grid / cl) : region;
  uint8_t* src; unsigned long long* d_cyc;
  cudaMalloc(&src, total + slot_bytes); cudaMemset(src, 1, total + slot_bytes); cudaMalloc(&d_cyc, 8);
  const size_t smem = 1024 + (size_t)slots * slot_bytes;
  void (*kern)(const uint8_t*, size_t, size_t, int, uint32_t, int, unsigned long long*) = cl == 1 ? k_fill<1> : (cl == 2 ? k_fill<2> : k_fill<4>);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem > 120 * 1024 ? smem : 120 * 1024;   // one CTA per SM
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, (const uint8_t*)src, region, stride, slots, slot_bytes, iters, d_cyc);
    cudaEventRecord(e1);
    cudaError_t se = cudaDeviceSynchronize();
    if (le != cudaSuccess || se != cudaSuccess) { printf("error: %s / %s\n", cudaGetErrorString(le), cudaGetErrorString(se)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long cyc; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    const double per_sm = (double)iters * slot_bytes;
    if (rep == 2)
      printf("cluster=%d slots=%d slot=%uKB region=%zuKB %s: %.1f GB/s per SM, %.2f TB/s chip-wide delivered, %.1f B/clk/SM (%.3f ms, %llu clk)\n", cl, slots,
             slot_bytes / 1024, region / 1024, distinct ? "distinct regions" : "same region", per_sm / (ms * 1e-3) / 1e9, per_sm * grid / (ms * 1e-3) / 1e12,
             per_sm / (double)cyc, ms, cyc);
  }
  return 0;
}
