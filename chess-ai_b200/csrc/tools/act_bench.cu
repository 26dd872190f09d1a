// act_bench.cu — what does the forward kernel's activation epilogue cost in isolation?
// Each warp holds kMaxNBH accumulator tiles in registers like mlp_forward_kernel and runs act_all(nbh) `iters` times
// (outputs are rescaled and fed back so nothing folds); 1..4 warps per SM sub-partition. Prints cycles per act_all call.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr act_bench.cu -o act_bench
#include <cstdio>
#include <vector>

#include "../mlp_stream_kernel.cuh"

namespace cab {
void set_error(const char *, ...) {}
int cuda_fail(cudaError_t, const char *, const char *, int) { return 1; }
}  // namespace cab
using namespace cab;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__global__ void __launch_bounds__(512, 1) act_kernel(const double *in, double *out, long long *cyc, int nbh, int iters, int mode) {
  __shared__ double tbl[kExpTblSize];
  for (int i = threadIdx.x; i < kExpTblSize; i += blockDim.x) tbl[i] = c_exp2_tbl[i] * 0.125;
  __syncthreads();
  double acc[kMaxNBH][2];
#pragma unroll
  for (int j = 0; j < kMaxNBH; ++j) {
    acc[j][0] = in[(threadIdx.x * kMaxNBH + j) * 2];
    acc[j][1] = in[(threadIdx.x * kMaxNBH + j) * 2 + 1];
  }
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) act_all(nbh, acc, tbl);
    else {   // the plain DFMA mix of the same length, no MUFU / LDS / integer work: the pipe-only floor
#pragma unroll
      for (int j = 0; j < kMaxNBH; ++j)
        if (j < nbh) {
#pragma unroll
          for (int k = 0; k < 13; ++k) { acc[j][0] = fma(acc[j][0], 1.0000001, 1e-9); acc[j][1] = fma(acc[j][1], 0.9999999, 1e-9); }
        }
    }
#pragma unroll
    for (int j = 0; j < kMaxNBH; ++j) { acc[j][0] = acc[j][0] * 3.1 + 0.37 * j; acc[j][1] = acc[j][1] * -2.9 + 0.11; }   // back into the active range
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < kMaxNBH; ++j) s += acc[j][0] + acc[j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = t1 - t0;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  std::vector<double> h(512 * kMaxNBH * 2);
  for (size_t i = 0; i < h.size(); ++i) h[i] = ((double) ((i * 2654435761u) % 2000) - 1000.0) / 37.0;
  double *d_in, *d_out;
  long long *d_cyc;
  CK(cudaMalloc(&d_in, h.size() * 8));
  CK(cudaMemcpy(d_in, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_out, (size_t) sms * 512 * 8));
  CK(cudaMalloc(&d_cyc, (size_t) sms * 16 * 8));
  const int iters = 200;
  for (int mode = 0; mode < 2; ++mode)
    for (int nbh : {15, 9, 7, 4, 1})
      for (int warps : {4, 8, 12, 16}) {
        act_kernel<<<sms, warps * 32>>>(d_in, d_out, d_cyc, nbh, iters, mode);
        CK(cudaDeviceSynchronize());
        std::vector<long long> c(sms * warps);
        CK(cudaMemcpy(c.data(), d_cyc, c.size() * 8, cudaMemcpyDeviceToHost));
        double avg = 0;
        for (auto v : c) avg += (double) v / c.size();
        const double per_call = avg / iters;
        // fp64-pipe floor: warps/4 per SMSP x 2 nbh activations x 13 operations x 2 cycles (+ the 2 rescale ops per value)
        const double floor = (warps / 4.0) * 2 * nbh * 13 * 2, fb = (warps / 4.0) * 2 * kMaxNBH * 1 * 2;
        printf("{\"mode\": \"%s\", \"nbh\": %d, \"warps_per_sm\": %d, \"cycles_per_call\": %.0f, \"pipe_floor\": %.0f, \"ratio\": %.2f}\n",
               mode == 0 ? "act_all" : "dfma_only", nbh, warps, per_call, floor + fb, per_call / (floor + fb));
      }
  return 0;
}
