// mufu_bench.cu — issue cost of the special-function instructions the fused fp16 evaluator's epilogue can be built from
// (MUFU.TANH, MUFU.EX2, MUFU.RCP, F2FP pack), per warp instruction and SM sub-partition, at 1 / 2 / 4 warps per SMSP.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu && ./mufu_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(float *out, long long *cycles, int iters) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.001f * (threadIdx.x + 37 * i) - 0.3f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 3) { unsigned u = __float_as_uint(v[i]); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u)); v[i] = __uint_as_float(u); }
      if (OP == 4) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i])); v[i] += 1.0f; asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i])); }
      if (OP == 5) v[i] = fmaf(v[i], 1.0001f, 0.5f);
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  float *out; long long *cyc;
  cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 1 << 16);
  const char *names[] = {"tanh.approx.f32", "ex2.approx.f32", "rcp.approx.f32", "tanh.approx.f16x2", "ex2 + add + rcp", "fma (reference)"};
  const int iters = 2000;
  for (int op = 0; op < 6; ++op)
    for (int warps : {4, 8, 16}) {      // per CTA = 1, 2, 4 per SMSP; one CTA per SM
      for (int rep = 0; rep < 2; ++rep) {
        switch (op) {
          case 0: k<0><<<148, warps * 32>>>(out, cyc, iters); break;
          case 1: k<1><<<148, warps * 32>>>(out, cyc, iters); break;
          case 2: k<2><<<148, warps * 32>>>(out, cyc, iters); break;
          case 3: k<3><<<148, warps * 32>>>(out, cyc, iters); break;
          case 4: k<4><<<148, warps * 32>>>(out, cyc, iters); break;
          case 5: k<5><<<148, warps * 32>>>(out, cyc, iters); break;
        }
        cudaDeviceSynchronize();
      }
      long long h[148];
      cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < 148; ++i) avg += h[i];
      avg /= 148;
      const double per_smsp_instr = avg / ((double) iters * 16 * (warps / 4));
      printf("{\"op\": \"%s\", \"warps_per_smsp\": %d, \"cycles_per_warp_instruction_per_smsp\": %.2f}\n", names[op], warps / 4, per_smsp_instr);
    }
  return 0;
}
