// mma_bench.cu — time per tcgen05.mma.cta_group::1.kind::f16 (M = 128, K = 16) as the fused fp16 evaluator issues them:
// A operand from tensor memory (.ts) or shared memory (.ss), B from a K-major SWIZZLE_128B shared-memory image, N in
// {112, 144, 240, 256}; back to back, or with a tcgen05.fence::after_thread_sync + mbarrier test between them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu && ./mma_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t) ((addr & 0x3FFFF) >> 4);
  d |= (uint64_t) (1024 >> 4) << 32;
  d |= (uint64_t) 1 << 46;
  d |= (uint64_t) 2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc(int n, int m = 128) { return (1u << 4) | ((uint32_t) (n >> 3) << 17) | ((uint32_t) (m >> 4) << 24); }
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
  } while (!ok);
}

// mode 0: .ts back to back; 1: .ts with fence + satisfied barrier wait before each; 2: .ss back to back; 3: .ss with M = 64;
// 4: .ts, four per loop iteration with descriptors advanced by constants (the GEMM kernel's issue pattern)
__global__ void k(int mode, int n, int n_mma, int reps, long long *cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ uint64_t bar, bar2;
  uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t) 1023);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(base)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar2)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar2)) : "memory");   // bar2: phase 0 complete
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t id = idesc(n, mode == 3 ? 64 : 128);
    const uint32_t wb = s32(base), ab = s32(base + 64 * 1024);
    long long total = 0;
    uint32_t ph = 0;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      if (mode == 4) {
        for (int kk = 0; kk < n_mma; kk += 4) {
          const uint64_t b = desc(wb + (uint32_t) ((kk >> 2) & 3) * (uint32_t) n * 128u);
          const uint32_t a = tmem + (uint32_t) (kk & 15) * 8u;
#pragma unroll
          for (int j = 0; j < 4; ++j) mma_ts(tmem + 256, a + 8u * j, b + (uint64_t) (2 * j), id, (kk != 0 || j != 0) ? 1u : 0u);
        }
      } else
      for (int kk = 0; kk < n_mma; ++kk) {
        const int k = kk & 15;
        const uint64_t b = desc(wb + (uint32_t) (k >> 2) * (uint32_t) n * 128u) + (uint64_t) ((k & 3) * 2);
        if (mode == 1) { mbar_wait(&bar2, 0); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
        if (mode == 2 || mode == 3) mma_ss(tmem + 256, desc(ab + (uint32_t) (k >> 2) * 128u * 128u) + (uint64_t) ((k & 3) * 2), b, id, kk != 0);
        else mma_ts(tmem + 256, tmem + (uint32_t) k * 8u, b, id, kk != 0);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
      mbar_wait(&bar, ph);
      ph ^= 1;
      total += clock64() - t0;
    }
    cycles[blockIdx.x] = total;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
  long long *cyc;
  cudaMalloc(&cyc, 1 << 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char *names[] = {"A in TMEM (.ts), back to back", "A in TMEM (.ts), barrier test + fence before each", "A in smem (.ss), back to back",
                         "A in smem (.ss), M = 64", "A in TMEM (.ts), 4 per iteration, constant descriptor steps"};
  for (int mode = 0; mode < 5; ++mode)
    for (int n : {16, 112, 256})
      for (int n_mma : {4, 16, 64}) {
        const int reps = 200;
        k<<<148, 128, 200 * 1024>>>(mode, n, n_mma, reps, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[148];
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < 148; ++i) avg += h[i];
        avg /= 148.0 * reps;
        printf("{\"mode\": \"%s\", \"N\": %d, \"mmas_per_commit\": %d, \"cycles_issue_to_commit\": %.0f, \"cycles_per_mma\": %.1f, \"floor_128N_over_256\": %.0f}\n",
               names[mode], n, n_mma, avg, avg / n_mma, 128.0 * n / 256.0);
      }
  return 0;
}
