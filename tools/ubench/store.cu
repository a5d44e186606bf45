// Micro-benchmark: DRAM write throughput of (0) plain 16-byte st.global, (1) st.global.cs (streaming), (2) shared-memory staging +
// cp.async.bulk (TMA) stores of `chunk` bytes, (3) 16-byte st.global with a read stream of equal size (copy).
//   ./store <mode> <MB> <chunk KB> <blocks per SM>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256) k_st(uint4* dst, const uint4* src, size_t n16, int mode) {
  const size_t stride = (size_t)gridDim.x * 256;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n16; i += stride) {
    uint4 v = make_uint4((uint32_t)i, 1u, 2u, 3u);
    if (mode == 3) v = __ldg(src + i);
    if (mode == 1) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else dst[i] = v;
  }
}

__global__ void __launch_bounds__(256) k_bulk(uint8_t* dst, size_t bytes, uint32_t chunk) {
  extern __shared__ __align__(128) uint8_t sm[];                 // two staging buffers of `chunk` bytes
  const size_t nchunks = bytes / chunk;
  int buf = 0;
  for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x, buf ^= 1) {
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store that last used this buffer has read it
    __syncthreads();
    uint4* s = reinterpret_cast<uint4*>(sm + (size_t)buf * chunk);
    for (uint32_t i = threadIdx.x; i < chunk / 16; i += 256) s[i] = make_uint4((uint32_t)c, i, 2u, 3u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * chunk), "r"(smem_u32(s)), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const size_t bytes = (size_t)(argc > 2 ? atoi(argv[2]) : 640) << 20;
  const uint32_t chunk = (argc > 3 ? atoi(argv[3]) : 8) * 1024u;
  const int bps = argc > 4 ? atoi(argv[4]) : 8;
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint8_t *dst, *src; cudaMalloc(&dst, bytes); cudaMalloc(&src, bytes); cudaMemset(src, 1, bytes);
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    if (mode == 2) k_bulk<<<sms * bps, 256, 2 * chunk>>>(dst, bytes, chunk);
    else k_st<<<sms * bps, 256>>>((uint4*)dst, (const uint4*)src, bytes / 16, mode);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const char* names[] = {"st.global.v4", "st.global.cs.v4", "smem + cp.async.bulk store", "copy (ld + st.global.v4)"};
  printf("%-28s %4zu MB chunk %3u KB, %d blocks/SM: %.3f ms  %.2f TB/s written%s\n", names[mode], bytes >> 20, chunk / 1024, bps, best,
         bytes / (best * 1e-3) / 1e12, mode == 3 ? " (+ the same read)" : "");
  return 0;
}
