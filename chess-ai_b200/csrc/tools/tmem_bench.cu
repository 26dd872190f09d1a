// tmem_bench.cu — cost of tensor-memory loads / stores as the fused fp16 evaluator's epilogue issues them
// (tcgen05.ld / st .32x32b), alone and mixed with MUFU.TANH, per warp instruction at 1 / 2 / 4 warps per lane quarter.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bench tmem_bench.cu && ./tmem_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MODE 0: ld16 + wait per iteration; 1: 4 x ld16 then one wait; 2: st8 + wait; 3: ld16 + wait + 16 tanh; 4: 16 tanh only;
// 5: ld16 + wait + 16 tanh + st8 + wait::st + fence::before_thread_sync (the evaluator's chunk without the barrier traffic)
template <int MODE>
__global__ void k(float *out, long long *cycles, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t) __cvta_generic_to_shared(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t) ((warp & 3) * 32) << 16);
  uint32_t r[4][16];
  uint32_t p[8] = {1, 2, 3, 4, 5, 6, 7, 8};
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 16; ++i) r[j][i] = 0;
  st8(base, p); wait_st();
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = (uint32_t) ((it * 16 + (warp >> 2) * 64) & 255);
    if (MODE == 0 || MODE == 3 || MODE == 5) { ld16(base + col, r[0]); wait_ld(); }
    if (MODE == 1) { ld16(base + col, r[0]); ld16(base + col + 16, r[1]); ld16(base + col + 32, r[2]); ld16(base + col + 48, r[3]); wait_ld(); }
    if (MODE == 2) { st8(base + col, p); wait_st(); }
    if (MODE == 3 || MODE == 4 || MODE == 5) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { float f = __uint_as_float(r[0][i]) + 0.001f * i; asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f)); acc += f; if (MODE == 5 && i < 8) p[i] = __float_as_uint(f); }
    }
    if (MODE == 5) { st8(base + 256 + col / 2, p); wait_st(); asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
    if (MODE != 4) acc += __uint_as_float(r[0][3] ^ r[1][5] ^ r[2][7] ^ r[3][9]);
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot));
}

int main() {
  float *out; long long *cyc;
  cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 1 << 16);
  const char *names[] = {"ld.x16 + wait::ld", "4 x ld.x16 + one wait::ld (per x16)", "st.x8 + wait::st", "ld.x16 + wait + 16 tanh", "16 tanh",
                         "ld.x16 + wait + 16 tanh + st.x8 + wait::st + fence"};
  const int iters = 4000;
  for (int mode = 0; mode < 6; ++mode)
    for (int warps : {4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        switch (mode) {
          case 0: k<0><<<148, warps * 32>>>(out, cyc, iters); break;
          case 1: k<1><<<148, warps * 32>>>(out, cyc, iters); break;
          case 2: k<2><<<148, warps * 32>>>(out, cyc, iters); break;
          case 3: k<3><<<148, warps * 32>>>(out, cyc, iters); break;
          case 4: k<4><<<148, warps * 32>>>(out, cyc, iters); break;
          case 5: k<5><<<148, warps * 32>>>(out, cyc, iters); break;
        }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long h[148];
      cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < 148; ++i) avg += h[i];
      avg /= 148;
      const double per = avg / ((double) iters * (mode == 1 ? 4 : 1) * (warps / 4));
      printf("{\"op\": \"%s\", \"warps_per_lane_quarter\": %d, \"cycles_per_16_columns_per_quarter\": %.1f}\n", names[mode], warps / 4, per);
    }
  return 0;
}
