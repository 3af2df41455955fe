// Microbenchmark (development): what does ONE cp.async.bulk (global -> shared, mbarrier completion) cost the issuing
// thread, and how many of them does an SM retire per cycle?  nvcc -arch=sm_100a -O3 -o tma_issue tma_issue.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const uint32_t *src, size_t stride_words, uint32_t bytes, int nops, int ctas_per_sm, uint64_t *out) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t *g = src + (size_t)blockIdx.x * stride_words;
    uint32_t parity = 0;
    long long t_issue = 0, t_done = 0;
    for (int rep = 0; rep < 8; rep++) {
      long long t0 = clock64();
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes * nops) : "memory");
      for (int i = 0; i < nops; i++) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(sm + (size_t)(i & 1) * bytes)),
                     "l"(g + (size_t)(rep * nops + i) * (bytes / 4)), "r"(bytes), "r"(smem_u32(&bar))
                     : "memory");
      }
      long long t1 = clock64();
      asm volatile(
          "{\n\t.reg .pred P1;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(
              smem_u32(&bar)),
          "r"(parity)
          : "memory");
      parity ^= 1;
      long long t2 = clock64();
      if (rep >= 2) { t_issue += t1 - t0; t_done += t2 - t0; }
    }
    out[2 * blockIdx.x] = t_issue / 6;
    out[2 * blockIdx.x + 1] = t_done / 6;
  }
}
int main() {
  int sms = 148;
  size_t words = (size_t)1 << 28;  // 1 GiB
  uint32_t *src; cudaMalloc(&src, words * 4); cudaMemset(src, 1, words * 4);
  uint64_t *out; cudaMallocManaged(&out, 148 * 8 * 2 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int cps : {1, 3}) for (uint32_t bytes : {512u, 4096u, 8192u, 16384u}) for (int nops : {1, 4, 12}) {
    int grid = sms * cps;
    size_t stride = (size_t)8 * nops * (bytes / 4);
    k<<<grid, 32, 2 * bytes + 64, 0>>>(src, stride, bytes, nops, cps, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    double a = 0, b = 0; for (int i = 0; i < grid; i++) { a += out[2 * i]; b += out[2 * i + 1]; }
    printf("ctas/SM %d bytes %6u nops %2d : issue %7.0f cyc (%.0f per op)  done %7.0f cyc  -> %.1f B/cyc/SM\n", cps, bytes, nops,
           a / grid, a / grid / nops, b / grid, (double)bytes * nops * cps / (b / grid));
  }
  return 0;
}
