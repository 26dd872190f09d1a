// tma_stream.cu — how fast can ONE CTA per SM stream the packed weight stream (541 KB, the same bytes for every CTA)
// from L2 into a shared-memory ring with 1-D TMA bulk copies? (The forward kernel's per-pass floor.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tma_stream.cu -o tma_stream
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(32, 1) stream_kernel(const uint8_t *src, long total, int chunk, int slots, int passes, long per_cta_stride,
                                                        unsigned long long *cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + (size_t) chunk * slots);
  src += (long) blockIdx.x * per_cta_stride;
  if (threadIdx.x == 0) {
    for (int s = 0; s < slots; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const long n_chunks = total / chunk * passes, per_pass = total / chunk;
    const long long t0 = clock64();
    for (long i = 0; i < n_chunks + slots; ++i) {
      const int s = (int) (i % slots);
      if (i >= slots) {   // wait for the copy issued `slots` ago into this slot
        const uint32_t parity = (uint32_t) (((i - slots) / slots) & 1);
        uint32_t ok;
        do {
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&bar[s])), "r"(parity) : "memory");
        } while (!ok);
      }
      if (i < n_chunks) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[s])), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s32(smem + (size_t) s * chunk)), "l"(src + (i % per_pass) * chunk), "r"(chunk), "r"(s32(&bar[s])) : "memory");
      }
    }
    cycles[blockIdx.x] = (unsigned long long) (clock64() - t0);
  }
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  const long total = 13 * 53248 / 1024 * 1024;   // ~ the packed forward stream of the default net
  uint8_t *src; CK(cudaMalloc(&src, total * (long) sms)); CK(cudaMemset(src, 1, total * (long) sms));
  unsigned long long *cyc; CK(cudaMalloc(&cyc, sms * sizeof(unsigned long long)));
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  const int passes = 64;
  for (int shared_src = 1; shared_src >= 0; --shared_src)
    for (int grid : {sms, 64, 8})
      for (auto cfg : {std::pair<int, int>{53248, 2}, {53248, 4}, {26624, 4}, {26624, 8}, {13312, 8}, {13312, 16}}) {
        const int chunk = cfg.first, slots = cfg.second;
        const size_t smem = (size_t) chunk * slots + 256;
        if (smem > 220 * 1024) continue;
        const long tot = total / chunk * chunk;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        stream_kernel<<<grid, 32, smem>>>(src, tot, chunk, slots, 2, shared_src ? 0 : total, cyc);   // warm L2
        cudaEventRecord(e0);
        stream_kernel<<<grid, 32, smem>>>(src, tot, chunk, slots, passes, shared_src ? 0 : total, cyc);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double) tot * passes;
        printf("{\"bench\": \"tma_stream\", \"same_bytes_for_all_ctas\": %d, \"ctas\": %d, \"chunk\": %d, \"slots\": %d, \"us_per_pass\": %.2f, "
               "\"GBps_per_sm\": %.1f, \"TBps_total\": %.2f, \"bytes_per_clk_per_sm\": %.1f}\n", shared_src, grid, chunk, slots, ms * 1e3 / passes,
               bytes / (ms * 1e-3) * 1e-9, bytes * grid / (ms * 1e-3) * 1e-12, bytes / (ms * 1e-3) / (p.clockRate * 1e3));
      }
  return 0;
}
