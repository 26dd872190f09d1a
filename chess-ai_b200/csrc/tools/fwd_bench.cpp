// fwd_bench.cpp — native timing harness of the fp64 forward path through the C ABI (no Python, starts in a second):
// used for kernel-variant sweeps (the library reads CAB_FWD_* from the environment once per process).
//   fwd_bench <weights.f64> [n_boards] [batch] [streams] [steps]
// Prints one JSON line: overlapped rate over `streams` streams, per-launch time on ONE stream (launches back to back),
// and the latency of one launch issued alone (event-timed, synchronised between launches).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../../include/chessai_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "CUDA %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
#define CAB(x) do { if ((x) != CAB_OK) { std::fprintf(stderr, "%s (line %d)\n", cab_last_error(), __LINE__); return 1; } } while (0)

int main(int argc, char **argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s weights.f64 [n_boards] [batch] [streams] [steps]\n", argv[0]); return 2; }
  const long n = argc > 2 ? std::atol(argv[2]) : 1L << 20;
  const long batch = argc > 3 ? std::atol(argv[3]) : 4096;
  const int n_streams = argc > 4 ? std::atoi(argv[4]) : 4;
  const int steps = argc > 5 ? std::atoi(argv[5]) : 6;
  cab_net *net = nullptr;
  CAB(cab_net_create(nullptr, 0, 0, &net));
  long nw = 0;
  CAB(cab_net_num_weights(net, &nw));
  std::vector<double> w(nw);
  FILE *f = std::fopen(argv[1], "rb");
  if (!f || std::fread(w.data(), sizeof(double), nw, f) != (size_t) nw) { std::fprintf(stderr, "cannot read %ld weights from %s\n", nw, argv[1]); return 1; }
  std::fclose(f);
  CAB(cab_net_set_weights(net, w.data(), nw));
  // boards: ~half the squares empty, the rest uniform codes (timing only; parity is pytest's job)
  const int sets = 4;
  std::vector<int8_t> h((size_t) n * 64);
  std::mt19937 rng(12345);
  int8_t *d_codes[sets];
  for (auto &c : h) { const unsigned r = rng(); c = (r & 1) ? 0 : (int8_t) ((int) ((r >> 1) % 19) - 9); }
  for (int s = 0; s < sets; ++s) {   // the same boards rotated by s * 1/8 of the set: distinct addresses, 4 x 64 MiB > L2
    const size_t rot = (size_t) n * 64 / 8 * s;
    CK(cudaMalloc(&d_codes[s], (size_t) n * 64));
    CK(cudaMemcpy(d_codes[s], h.data() + rot, (size_t) n * 64 - rot, cudaMemcpyHostToDevice));
    if (rot) CK(cudaMemcpy(d_codes[s] + ((size_t) n * 64 - rot), h.data(), rot, cudaMemcpyHostToDevice));
  }
  double *d_out;
  CK(cudaMalloc(&d_out, n * sizeof(double)));
  std::vector<cudaStream_t> streams(n_streams);
  for (auto &s : streams) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  cudaStream_t main_s;
  CK(cudaStreamCreateWithFlags(&main_s, cudaStreamNonBlocking));
  std::vector<cudaEvent_t> joins(n_streams);
  for (auto &e : joins) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  cudaEvent_t fork;
  CK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
  auto step = [&](int i) -> int {
    CK(cudaEventRecord(fork, main_s));
    for (auto &s : streams) CK(cudaStreamWaitEvent(s, fork, 0));
    CAB(cab_eval_dev_batches(net, d_codes[i % sets], n, batch, d_out, (void *const *) streams.data(), n_streams));
    for (int k = 0; k < n_streams; ++k) { CK(cudaEventRecord(joins[k], streams[k])); CK(cudaStreamWaitEvent(main_s, joins[k], 0)); }
    return 0;
  };
  for (int i = 0; i < 3; ++i) if (step(i)) return 1;
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0, main_s));
  for (int i = 0; i < steps; ++i) if (step(i)) return 1;
  CK(cudaEventRecord(e1, main_s));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double rate = (double) n * steps / (ms * 1e-3);
  // one stream, launches back to back
  const long n1 = std::min<long>(n, 64 * batch);
  void *one[1] = {streams[0]};
  CAB(cab_eval_dev_batches(net, d_codes[0], n1, batch, d_out, one, 1));
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0, streams[0]));
  CAB(cab_eval_dev_batches(net, d_codes[1], n1, batch, d_out, one, 1));
  CK(cudaEventRecord(e1, streams[0]));
  CK(cudaDeviceSynchronize());
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double serial_us = ms * 1e3 / (double) ((n1 + batch - 1) / batch);
  // one launch alone
  double alone_us = 0;
  const int reps = 50;
  for (int r = 0; r < reps + 5; ++r) {
    CK(cudaEventRecord(e0, streams[0]));
    CAB(cab_eval_dev(net, d_codes[r % sets] + (size_t) (r % 8) * batch * 64, batch, d_out, streams[0]));
    CK(cudaEventRecord(e1, streams[0]));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r >= 5) alone_us += ms * 1e3 / reps;
  }
  double peak = 0;
  CAB(cab_measure_pipe_peak(0, 0, &peak));
  const double flop = 128432.0;
  std::printf("{\"n_boards\": %ld, \"batch\": %ld, \"streams\": %d, \"steps\": %d, \"evals_per_s\": %.4g, \"tflops\": %.3f, \"dmma_peak\": %.2f, "
              "\"frac\": %.4f, \"launch_us_single_stream\": %.2f, \"launch_us_alone\": %.2f, \"frac_single_launch\": %.4f, \"env\": {\"lag\": \"%s\", \"shift\": \"%s\", "
              "\"pairs\": \"%s\", \"stage_kb\": \"%s\", \"stages\": \"%s\"}}\n",
              n, batch, n_streams, steps, rate, rate * flop * 1e-12, peak, rate * flop * 1e-12 / peak, serial_us, alone_us,
              flop * batch / (alone_us * 1e-6) * 1e-12 / peak,
              getenv("CAB_FWD_LAG") ? getenv("CAB_FWD_LAG") : "", getenv("CAB_FWD_TEAM_SHIFT") ? getenv("CAB_FWD_TEAM_SHIFT") : "",
              getenv("CAB_FWD_PAIRS") ? getenv("CAB_FWD_PAIRS") : "", getenv("CAB_STAGE_KB") ? getenv("CAB_STAGE_KB") : "",
              getenv("CAB_FWD_STAGES") ? getenv("CAB_FWD_STAGES") : "");
  cab_net_destroy(net);
  return 0;
}
