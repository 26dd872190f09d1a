// pipe_peaks.cu — standalone microbenchmark for the roofline denominators that
// MEASURED_PEAKS.json does not carry ("fp64 and fp32 pipe peaks ...
// must be measured on the box with the same best-of-10 CUDA-event method").
//
// Measures, on the current device, best-of-10 with CUDA events:
//   dfma      : fp64 FMA pipe (independent DFMA chains)
//   dmma      : fp64 tensor pipe (mma.sync m8n8k4 f64 → SASS DMMA.8x8x4)
//   dmma+dfma : both issued from the same warps (are the pipes shared?)
//   ffma      : fp32 FMA pipe
//   act       : cost of the reference activation f(x)=4*(2/(1+exp(-2*(x/4)))-1) in fp64
//   lds       : shared-memory wavefront behaviour of the fragment-load patterns we use
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o pipe_peaks pipe_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kThreads = 256;

template <int NACC>
__global__ void __launch_bounds__(kThreads) dfma_kernel(double* out, int iters, double seed) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = seed + i + threadIdx.x;
  double a = seed * 1.0000001, b = seed * 0.5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(kThreads) ffma_kernel(float* out, int iters, float seed) {
  float acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = seed + i + threadIdx.x;
  float a = seed * 1.0000001f, b = seed * 0.5f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// NT independent accumulator tiles per warp; NF extra DFMA per DMMA (0 = pure DMMA)
template <int NT, int NF>
__global__ void __launch_bounds__(kThreads) dmma_kernel(double* out, int iters, double seed) {
  double c[NT][2];
  double f[NF > 0 ? NF * 4 : 1];
#pragma unroll
  for (int i = 0; i < NT; ++i) { c[i][0] = seed + i; c[i][1] = seed - i; }
#pragma unroll
  for (int i = 0; i < (NF > 0 ? NF * 4 : 1); ++i) f[i] = seed + i;
  double a = seed * 1e-3 + threadIdx.x * 1e-9, b = seed * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      dmma884(c[i][0], c[i][1], a, b);
      if (NF > 0) {
#pragma unroll
        for (int j = 0; j < NF; ++j) f[(i * NF + j) % (NF * 4)] = fma(f[(i * NF + j) % (NF * 4)], a, b);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < (NF > 0 ? NF * 4 : 1); ++i) s += f[i];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// reference activation as written (network.h:83-94 of the reference): 4*(2/(1+exp(-2*(x/4)))-1)
__device__ __forceinline__ double act_ref(double x) {
  return 4.0 * (2.0 / (1.0 + exp(-2.0 * (x / 4.0))) - 1.0);
}
template <int NACC>
__global__ void __launch_bounds__(kThreads) act_kernel(double* out, int iters, double seed) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = seed * (i + 1) + threadIdx.x * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = act_ref(acc[i] * 1.7 - 0.1);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// LDS pattern probes: cycles per warp-instruction, one warp per SM-quadrant... measured with clock64
// mode 0: LDS.64, every lane a distinct consecutive double (256 B/warp)
// mode 1: LDS.128, every lane distinct consecutive 16 B (512 B/warp)
// mode 2: LDS.128, 4 distinct 16 B chunks broadcast over 8 lanes each
// mode 3: LDS.128, all lanes the same 16 B
// mode 4: LDS.64, 8 distinct doubles broadcast over 4 lanes each
template <int MODE>
__global__ void __launch_bounds__(kThreads) lds_kernel(double* out, long long* cyc, int iters) {
  __shared__ __align__(16) double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int base;
  if (MODE == 0) base = lane; else if (MODE == 1) base = lane * 2; else if (MODE == 2) base = (lane >> 3) * 2;
  else if (MODE == 3) base = 0; else base = lane >> 2;
  base += warp * 64;
  double s0 = 0, s1 = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      int idx = (base + u * 128 + it * 2) & 4095;
      if (MODE == 0 || MODE == 4) { s0 += sm[idx]; }
      else { double2 v = *reinterpret_cast<const double2*>(&sm[idx & ~1]); s0 += v.x; s1 += v.y; }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (s0 + s1 == 12345.678) out[threadIdx.x] = s0 + s1;
}

// accuracy of the MUFU.RCP64H seed (rcp.approx.ftz.f64) and of 1..3 Newton steps, over d in [1, 2^k)
__global__ void rcp_probe_kernel(double* maxerr) {
  double e0 = 0, e1 = 0, e2 = 0, e3 = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (1 << 22); i += gridDim.x * blockDim.x) {
    const double d = (1.0 + i * (1.0 / (1 << 22))) * (double) (1ull << (i % 40));
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    e0 = fmax(e0, fabs(fma(-d, y, 1.0)));
    double r = fma(-d, y, 1.0); y = fma(y, r, y);
    e1 = fmax(e1, fabs(fma(-d, y, 1.0)));
    r = fma(-d, y, 1.0); y = fma(y, r, y);
    e2 = fmax(e2, fabs(fma(-d, y, 1.0)));
    r = fma(-d, y, 1.0); y = fma(y, r, y);
    e3 = fmax(e3, fabs(fma(-d, y, 1.0)));
  }
  // max over the grid via atomicMax on the bit pattern (all values are non-negative)
  atomicMax((unsigned long long*) &maxerr[0], (unsigned long long) __double_as_longlong(e0));
  atomicMax((unsigned long long*) &maxerr[1], (unsigned long long) __double_as_longlong(e1));
  atomicMax((unsigned long long*) &maxerr[2], (unsigned long long) __double_as_longlong(e2));
  atomicMax((unsigned long long*) &maxerr[3], (unsigned long long) __double_as_longlong(e3));
}

template <typename F>
static double best_ms(F launch, int reps = 10) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"clock_khz\": %d}\n", p.name, sms, p.major, p.minor, p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double) * 1 << 22));
  const int iters = 4096;
  for (int bps : {1, 2, 4}) {
    int grid = sms * bps;
    {
      constexpr int NACC = 16;
      double ms = best_ms([&] { dfma_kernel<NACC><<<grid, kThreads>>>(out, iters, 1.0); });
      double fl = 2.0 * NACC * iters * (double)grid * kThreads;
      printf("{\"bench\": \"dfma\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", bps, ms, fl / ms * 1e-9);
    }
    {
      constexpr int NT = 16;
      double ms = best_ms([&] { dmma_kernel<NT, 0><<<grid, kThreads>>>(out, iters, 1.0); });
      double fl = 2.0 * 256 * NT * iters * (double)grid * (kThreads / 32);
      printf("{\"bench\": \"dmma884\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", bps, ms, fl / ms * 1e-9);
    }
    {
      constexpr int NT = 16, NF = 2;  // 2 DFMA (64 flop/warp-instr each... per thread 2 flop) per DMMA
      double ms = best_ms([&] { dmma_kernel<NT, NF><<<grid, kThreads>>>(out, iters, 1.0); });
      double fl_mma = 2.0 * 256 * NT * iters * (double)grid * (kThreads / 32);
      double fl_fma = 2.0 * NF * NT * iters * (double)grid * kThreads;
      printf("{\"bench\": \"dmma884+2dfma\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"tflops_mma\": %.3f, \"tflops_fma\": %.3f}\n",
             bps, ms, fl_mma / ms * 1e-9, fl_fma / ms * 1e-9);
    }
    {
      constexpr int NT = 16, NF = 8;  // equal flop on both pipes
      double ms = best_ms([&] { dmma_kernel<NT, NF><<<grid, kThreads>>>(out, iters, 1.0); });
      double fl_mma = 2.0 * 256 * NT * iters * (double)grid * (kThreads / 32);
      double fl_fma = 2.0 * NF * NT * iters * (double)grid * kThreads;
      printf("{\"bench\": \"dmma884+8dfma\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"tflops_mma\": %.3f, \"tflops_fma\": %.3f}\n",
             bps, ms, fl_mma / ms * 1e-9, fl_fma / ms * 1e-9);
    }
    {
      constexpr int NACC = 32;
      double ms = best_ms([&] { ffma_kernel<NACC><<<grid, kThreads>>>((float*)out, iters, 1.0f); });
      double fl = 2.0 * NACC * iters * (double)grid * kThreads;
      printf("{\"bench\": \"ffma\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", bps, ms, fl / ms * 1e-9);
    }
    {
      constexpr int NACC = 8;
      int it2 = 256;
      double ms = best_ms([&] { act_kernel<NACC><<<grid, kThreads>>>(out, it2, 0.37); });
      double n = (double)NACC * it2 * grid * kThreads;
      printf("{\"bench\": \"act_ref_fp64\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"gact_per_s\": %.3f}\n", bps, ms, n / ms * 1e-6);
    }
  }
  // dependent-chain latency of DMMA (1 warp, 1 tile)
  {
    double ms = best_ms([&] { dmma_kernel<1, 0><<<1, 32>>>(out, 1 << 16, 1.0); });
    printf("{\"bench\": \"dmma884_latency\", \"ns_per_dependent_mma\": %.3f}\n", ms * 1e6 / (1 << 16));
    double ms2 = best_ms([&] { dfma_kernel<1><<<1, 32>>>(out, 1 << 16, 1.0); });
    printf("{\"bench\": \"dfma_latency\", \"ns_per_dependent_fma\": %.3f}\n", ms2 * 1e6 / (1 << 16));
  }
  // per-warp DMMA issue rate with few warps: 1 warp/SMSP, 8 independent tiles
  {
    double* me; CK(cudaMallocManaged(&me, 4 * sizeof(double)));
    for (int i = 0; i < 4; ++i) me[i] = 0;
    rcp_probe_kernel<<<sms, 256>>>(me); CK(cudaDeviceSynchronize());
    printf("{\"bench\": \"rcp_approx_f64\", \"max_abs_1_minus_dy\": {\"seed\": %.3e, \"nr1\": %.3e, \"nr2\": %.3e, \"nr3\": %.3e}}\n", me[0], me[1], me[2], me[3]);
  }
  for (int warps : {1, 4, 8}) {
    double ms = best_ms([&] { dmma_kernel<16, 0><<<sms, warps * 32>>>(out, iters, 1.0); });
    double fl = 2.0 * 256 * 16 * iters * (double)sms * warps;
    printf("{\"bench\": \"dmma884_warps\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", warps, ms, fl / ms * 1e-9);
  }
  long long* cyc; CK(cudaMallocManaged(&cyc, sizeof(long long) * 1024));
  {
    const int li = 256;
    auto rep = [&](const char* name, int mode) {
      CK(cudaDeviceSynchronize());
      printf("{\"bench\": \"lds\", \"pattern\": \"%s\", \"mode\": %d, \"cycles_per_warp_instr_8warps\": %.3f}\n", name, mode,
             (double)cyc[0] / (li * 16.0 * 8.0));
    };
    lds_kernel<0><<<1, 256>>>(out, cyc, li); rep("lds64_distinct", 0);
    lds_kernel<1><<<1, 256>>>(out, cyc, li); rep("lds128_distinct", 1);
    lds_kernel<2><<<1, 256>>>(out, cyc, li); rep("lds128_bcast4x8", 2);
    lds_kernel<3><<<1, 256>>>(out, cyc, li); rep("lds128_bcast_all", 3);
    lds_kernel<4><<<1, 256>>>(out, cyc, li); rep("lds64_bcast8x4", 4);
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
