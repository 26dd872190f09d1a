// cab_leafq.cu — the MCTS leaf queue and K7, the prior softmax.
//
// The reference evaluates ONE board per node expansion inside Decider::prediction
// (/root/reference/src/mcts_network/decider.cpp:31-61, called from tree.cpp:295). With virtual loss the search keeps
// many playouts in flight, so expansions of many trees / games pile up here and are evaluated thousands per launch:
//   worker threads  cab_leafq_push(codes)        -> slot numbers (lock-free reservation, copy into pinned memory)
//   any thread      cab_leafq_flush()            -> H2D in pieces, cab_eval launches of `batch` boards round-robin
//                                                   over the queue's streams, D2H; results indexed by slot
// K7: priors of a node's children, tree.cpp:299-306:  w_i = exp(cm * logit_i),  prior_i = w_i / sum_j w_j
#include <atomic>
#include <cstring>

#include "cab_internal.cuh"

extern "C" int cab_eval_f16_dev(cab_net *net, const int8_t *codes_dev, long n, double *out_dev, void *stream);
namespace cab {

// one warp per segment (node); segments are short (a chess position has ~20-40 children)
__global__ void __launch_bounds__(256) prior_softmax_kernel(const double *__restrict__ logits, const int *__restrict__ seg_off,
                                                             int n_seg, const double *__restrict__ color_mult,
                                                             double *__restrict__ priors) {
  const int seg = (int) ((blockIdx.x * (long) blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (seg >= n_seg) return;
  const int lo = seg_off[seg], hi = seg_off[seg + 1];
  const double cm = color_mult[seg];
  double sum = 0.0;
  for (int i = lo + lane; i < hi; i += 32) sum += exp(cm * logits[i]);       // tree.cpp:302-303
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  for (int i = lo + lane; i < hi; i += 32) priors[i] = exp(cm * logits[i]) / sum;  // Node::expand, tree.cpp:92
}
}  // namespace cab

struct cab_leafq {
  cab_net *net = nullptr;
  long capacity = 0, batch = 4096;
  int n_streams = 4;
  int8_t *h_codes = nullptr;   // pinned [capacity][64]
  double *h_out = nullptr;     // pinned [capacity]
  int8_t *d_codes = nullptr;
  double *d_out = nullptr;
  std::vector<cudaStream_t> streams;
  std::vector<cudaEvent_t> events;
  std::atomic<long> head{0};
  long launches = 0, flushes = 0, evaluated = 0;
  bool f16 = false;            // evaluate with the reduced-precision fused tcgen05 kernel (cab_leafq_set_precision)
  long in_flight = -1;         // boards of the flush queued by cab_leafq_flush_async and not yet waited for
};

using namespace cab;

static int leafq_allocate(cab_leafq *q) {
  CAB_CUDA(cudaMallocHost(&q->h_codes, q->capacity * 64));
  CAB_CUDA(cudaMallocHost(&q->h_out, q->capacity * sizeof(double)));
  CAB_CUDA(cudaMalloc(&q->d_codes, q->capacity * 64));
  CAB_CUDA(cudaMalloc(&q->d_out, q->capacity * sizeof(double)));
  q->streams.assign(q->n_streams, nullptr);
  q->events.assign(q->n_streams, nullptr);
  for (int i = 0; i < q->n_streams; ++i) {
    CAB_CUDA(cudaStreamCreateWithFlags(&q->streams[i], cudaStreamNonBlocking));
    CAB_CUDA(cudaEventCreateWithFlags(&q->events[i], cudaEventDisableTiming));
  }
  return CAB_OK;
}

extern "C" {

int cab_leafq_destroy(cab_leafq *q);

int cab_leafq_create(cab_net *net, long capacity, long batch, int n_streams, cab_leafq **out) {
  if (!net || !out || capacity <= 0 || batch <= 0 || n_streams <= 0 || n_streams > 32) { set_error("bad argument"); return CAB_EINVAL; }
  *out = nullptr;
  CAB_CUDA(cudaSetDevice(net->device));
  cab_leafq *q = new cab_leafq();
  q->net = net; q->capacity = capacity; q->batch = batch; q->n_streams = n_streams;
  const int rc = leafq_allocate(q);
  if (rc != CAB_OK) { cab_leafq_destroy(q); return rc; }   // a partially built queue is released, not leaked
  *out = q;
  return CAB_OK;
}

int cab_leafq_destroy(cab_leafq *q) {
  if (!q) return CAB_OK;
  cudaSetDevice(q->net->device);
  for (auto s : q->streams) if (s) cudaStreamDestroy(s);
  for (auto e : q->events) if (e) cudaEventDestroy(e);
  cudaFreeHost(q->h_codes); cudaFreeHost(q->h_out); cudaFree(q->d_codes); cudaFree(q->d_out);
  delete q;
  return CAB_OK;
}

// Thread-safe among pushers. Reserves n consecutive slots (compare-and-swap: a reservation either fits entirely or
// leaves `head` untouched, so no pusher ever sees an inflated head or shares a slot) and copies the boards in;
// *first_slot is the ticket for the results. CAB_ESTATE when the queue is full (flush first).
// The copy happens AFTER the reservation: a flush must not overlap pushes that are still in progress.
int cab_leafq_push(cab_leafq *q, const int8_t *codes, long n, long *first_slot) {
  if (!q || !codes || n <= 0 || !first_slot) { set_error("bad argument"); return CAB_EINVAL; }
  long at = q->head.load(std::memory_order_relaxed);
  do {
    if (at + n > q->capacity) {
      set_error("leaf queue full (%ld + %ld > %ld): flush it", at, n, q->capacity);
      return CAB_ESTATE;
    }
  } while (!q->head.compare_exchange_weak(at, at + n, std::memory_order_acq_rel, std::memory_order_relaxed));
  memcpy(q->h_codes + at * 64, codes, (size_t) n * 64);
  *first_slot = at;
  return CAB_OK;
}

int cab_leafq_pending(cab_leafq *q, long *n) {
  if (!q || !n) { set_error("null argument"); return CAB_EINVAL; }
  *n = q->head.load();
  return CAB_OK;
}

// Queues the evaluation of slots [0, pending) on the queue's streams (H2D, launches, D2H) and returns without waiting.
// No push until cab_leafq_wait. A thread that owns two queues can walk its trees for one while the GPU works on the other.
int cab_leafq_flush_async(cab_leafq *q, long *n_evaluated) {
  if (!q) { set_error("null argument"); return CAB_EINVAL; }
  if (q->in_flight >= 0) { set_error("a flush of this queue is already in flight: cab_leafq_wait first"); return CAB_ESTATE; }
  const long n = q->head.load();
  if (n_evaluated) *n_evaluated = n;
  q->in_flight = n;
  if (n == 0) return CAB_OK;
  cab_net *net = q->net;
  CAB_CUDA(cudaSetDevice(net->device));
  if (q->f16) {           // reduced precision: one fused launch per stream piece
    const long piece = (n + q->n_streams - 1) / q->n_streams;
    int i = 0;
    for (long b0 = 0; b0 < n; b0 += piece, ++i) {
      const long nb = std::min(piece, n - b0);
      cudaStream_t s = q->streams[i];
      CAB_CUDA(cudaMemcpyAsync(q->d_codes + b0 * 64, q->h_codes + b0 * 64, nb * 64, cudaMemcpyHostToDevice, s));
      const int rc = cab_eval_f16_dev(net, q->d_codes + b0 * 64, nb, q->d_out + b0, s);   // (locks only to rebuild the image)
      if (rc != CAB_OK) return rc;
      CAB_CUDA(cudaMemcpyAsync(q->h_out + b0, q->d_out + b0, nb * sizeof(double), cudaMemcpyDeviceToHost, s));
      ++q->launches;
    }
    return CAB_OK;
  }
  {  // weights changed since the last evaluation (rare): re-pack once. The flag is only ever read under the net's lock
     // and cleared after the pack stream has drained, so no other thread's launch can overtake the pack kernels.
    std::lock_guard<std::mutex> lk(net->mu);
    const int rc = ensure_packed(net, q->streams[0], true);
    if (rc != CAB_OK) return rc;
  }
  int i = 0;
  for (long b0 = 0; b0 < n; b0 += q->batch, ++i) {
    const long nb = std::min(q->batch, n - b0);
    cudaStream_t s = q->streams[i % q->n_streams];
    CAB_CUDA(cudaMemcpyAsync(q->d_codes + b0 * 64, q->h_codes + b0 * 64, nb * 64, cudaMemcpyHostToDevice, s));
    // search threads wait for this flush: spread every launch over the whole device (latency geometry)
    int rc = launch_forward(net, q->d_codes + b0 * 64, nb, q->d_out + b0, nullptr, 0, s, nullptr, /*alone=*/true);
    if (rc != CAB_OK) return rc;
    CAB_CUDA(cudaMemcpyAsync(q->h_out + b0, q->d_out + b0, nb * sizeof(double), cudaMemcpyDeviceToHost, s));
    ++q->launches;
  }
  return CAB_OK;
}

// Waits for the flush in flight (if any); the results are then valid and the queue is empty again.
int cab_leafq_wait(cab_leafq *q) {
  if (!q) { set_error("null argument"); return CAB_EINVAL; }
  if (q->in_flight < 0) return CAB_OK;
  const long n = q->in_flight;
  q->in_flight = -1;
  if (n > 0) {
    CAB_CUDA(cudaSetDevice(q->net->device));
    for (auto s : q->streams) CAB_CUDA(cudaStreamSynchronize(s));
    ++q->flushes;
    q->evaluated += n;
  }
  q->head.store(0);
  return CAB_OK;
}

// Not concurrent with push. Evaluates slots [0, pending) and empties the queue.
int cab_leafq_flush(cab_leafq *q, long *n_evaluated) {
  const int rc = cab_leafq_flush_async(q, n_evaluated);
  if (rc != CAB_OK) { if (q) q->in_flight = -1; return rc; }
  return cab_leafq_wait(q);
}

// precision = 0: the fp64 reference-precision kernel (default); 1: the fused fp16 tcgen05 kernel (cab_eval_f16_*: nets
// shaped {64, N1, N2, N3, 1} only, reduced precision, no parity claim)
int cab_leafq_set_precision(cab_leafq *q, int precision) {
  if (!q || (precision != 0 && precision != 1)) { set_error("bad argument"); return CAB_EINVAL; }
  q->f16 = precision == 1;
  return CAB_OK;
}

// results of the last flush, indexed by slot; valid until the next push
int cab_leafq_results(cab_leafq *q, const double **results) {
  if (!q || !results) { set_error("null argument"); return CAB_EINVAL; }
  *results = q->h_out;
  return CAB_OK;
}

int cab_leafq_stats(cab_leafq *q, long *launches, long *flushes, long *evaluated) {
  if (!q) { set_error("null argument"); return CAB_EINVAL; }
  if (launches) *launches = q->launches;
  if (flushes) *flushes = q->flushes;
  if (evaluated) *evaluated = q->evaluated;
  return CAB_OK;
}

// K7 on host arrays: n_seg nodes, node s owns logits[seg_off[s] .. seg_off[s+1]) and colour multiplier color_mult[s]
// Stand-alone K7 for callers that keep their search on the host (inside cab_mcts_* it is fused into the walk kernel). The
// device scratch is one grow-only block per calling thread and device: no allocation in the steady state.
int cab_prior_softmax_host(int device, const double *logits, const int *seg_off, int n_seg, const double *color_mult,
                           double *priors_out) {
  if (n_seg < 0 || (n_seg > 0 && (!logits || !seg_off || !color_mult || !priors_out))) { set_error("bad argument"); return CAB_EINVAL; }
  if (n_seg == 0) return CAB_OK;
  CAB_CUDA(cudaSetDevice(device));
  const int total = seg_off[n_seg];
  if (total < 0) { set_error("seg_off must be non-decreasing"); return CAB_EINVAL; }
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t b_l = up((size_t) std::max(total, 1) * sizeof(double)), b_c = up((size_t) n_seg * sizeof(double)),
               b_o = up((size_t) (n_seg + 1) * sizeof(int));
  const size_t need = 2 * b_l + b_c + b_o;
  thread_local struct { int device = -1; size_t cap = 0; uint8_t *p = nullptr; } scratch;
  if (scratch.device != device || scratch.cap < need) {
    if (scratch.p != nullptr) { cudaSetDevice(scratch.device); cudaFree(scratch.p); cudaSetDevice(device); }
    scratch.p = nullptr; scratch.cap = 0; scratch.device = device;
    CAB_CUDA(cudaMalloc(&scratch.p, need + need / 2));
    scratch.cap = need + need / 2;
  }
  double *d_l = reinterpret_cast<double *>(scratch.p), *d_p = reinterpret_cast<double *>(scratch.p + b_l),
         *d_c = reinterpret_cast<double *>(scratch.p + 2 * b_l);
  int *d_o = reinterpret_cast<int *>(scratch.p + 2 * b_l + b_c);
  CAB_CUDA(cudaMemcpyAsync(d_l, logits, (size_t) total * sizeof(double), cudaMemcpyHostToDevice, cudaStreamPerThread));
  CAB_CUDA(cudaMemcpyAsync(d_c, color_mult, (size_t) n_seg * sizeof(double), cudaMemcpyHostToDevice, cudaStreamPerThread));
  CAB_CUDA(cudaMemcpyAsync(d_o, seg_off, (size_t) (n_seg + 1) * sizeof(int), cudaMemcpyHostToDevice, cudaStreamPerThread));
  prior_softmax_kernel<<<(n_seg * 32 + 255) / 256, 256, 0, cudaStreamPerThread>>>(d_l, d_o, n_seg, d_c, d_p);
  CAB_CUDA(cudaGetLastError());
  CAB_CUDA(cudaMemcpyAsync(priors_out, d_p, (size_t) total * sizeof(double), cudaMemcpyDeviceToHost, cudaStreamPerThread));
  CAB_CUDA(cudaStreamSynchronize(cudaStreamPerThread));
  return CAB_OK;
}

}  // extern "C"
