// cab_train.cu — training entry points of include/chessai_b200.h (Optimizer::optimize / dropout, train loop, DP).
#include <cstdio>

#include "cab_internal.cuh"

namespace cab {

// Optimizer::dropout ("weak dilution", /root/reference/src/mcts_network/network.cpp:281-292): every weight is
// replaced by U(-1,1) with probability `rate`. The reference walks the weights with a sequential mt19937
// (one draw per weight + one per hit); here weight w owns Philox4x32-10(key = seed, counter = (w, offset)):
// words 0,1 -> chance draw, words 2,3 -> replacement. Mirrors oracle/mlp_oracle.c:orc_dropout_philox bit for bit.
__global__ void __launch_bounds__(256) dropout_kernel(double *__restrict__ w, long n, double rate, uint64_t seed,
                                                       uint64_t offset, unsigned long long *__restrict__ hits) {
  unsigned long long local = 0;
  for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
    uint32_t r[4];
    philox4x32_10((uint32_t) i, (uint32_t) ((uint64_t) i >> 32), (uint32_t) offset, (uint32_t) (offset >> 32),
                  (uint32_t) seed, (uint32_t) (seed >> 32), r);
    if (u53(r[0], r[1]) < rate) {
      w[i] = 2.0 * u53(r[2], r[3]) + -1.0;
      ++local;
    }
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(hits, local);
}

}  // namespace cab

using namespace cab;

extern "C" {

int cab_dropout(cab_net *net, double rate, uint64_t seed, uint64_t offset, long *hits) {
  if (!net) { set_error("null argument"); return CAB_EINVAL; }
  CAB_CUDA(cudaSetDevice(net->device));
  unsigned long long *d_hits = nullptr;
  CAB_CUDA(cudaMalloc(&d_hits, sizeof(unsigned long long)));
  CAB_CUDA(cudaMemset(d_hits, 0, sizeof(unsigned long long)));
  dropout_kernel<<<(int) std::min<long>((net->n_weights + 255) / 256, 148L * 8), 256>>>(net->d_w, net->n_weights, rate,
                                                                                          seed, offset, d_hits);
  net->launches++;
  CAB_CUDA(cudaGetLastError());
  unsigned long long h = 0;
  CAB_CUDA(cudaMemcpy(&h, d_hits, sizeof h, cudaMemcpyDeviceToHost));
  cudaFree(d_hits);
  if (hits) *hits = (long) h;
  net->pack_dirty = true;
  return CAB_OK;
}

int cab_train_case(cab_net *, const int8_t *, double, double *, double, double, int *, double *, double *) {
  set_error("cab_train_case: not built yet");
  return CAB_ESTATE;
}
int cab_train_sequence(cab_net *, const int8_t *, const double *, long, double *, uint64_t, uint8_t *, int *) {
  set_error("cab_train_sequence: not built yet");
  return CAB_ESTATE;
}
int cab_train_batch_host(cab_net *, const int8_t *, const double *, long, double *, double, double, int *, double *,
                         double *) {
  set_error("cab_train_batch_host: not built yet");
  return CAB_ESTATE;
}
int cab_train_batch_dev(cab_net *, const int8_t *, const double *, long, double *, double, double, int *, double *,
                        double *, void *) {
  set_error("cab_train_batch_dev: not built yet");
  return CAB_ESTATE;
}
int cab_dp_unique_id(void *) { set_error("cab_dp: not built yet"); return CAB_ESTATE; }
int cab_dp_init(cab_net *, const void *, int, int) { set_error("cab_dp: not built yet"); return CAB_ESTATE; }
int cab_dp_shutdown(cab_net *) { return CAB_OK; }
int cab_grad_buffer(cab_net *, double **, long *) { set_error("cab_grad_buffer: not built yet"); return CAB_ESTATE; }

}  // extern "C"
