// cab_tc.cu — host side of the bf16 tcgen05 path (cab_eval_bf16_*): weight packing, TMA tensor maps, layer loop.
#include <cuda.h>

#include <cstdio>

#include "cab_internal.cuh"
#include "tc_gemm_bf16.cuh"
#include "mlp_tc_fused.cuh"

namespace cab {

struct TcState {
  std::vector<__nv_bfloat16 *> wt;  // per GEMM layer: Wt[N][K]
  std::vector<__nv_bfloat16 *> wkn; // per GEMM layer l >= 1: W[K][N] (operand of the dX GEMM), packed on first training use
  bool wkn_packed = false;
  float *w_last = nullptr;          // last layer (N = 1) in fp32
  // training buffers (tc_train_prepare): saved activations a_l and psi_l, row-major [B][width] and transposed [width][ld]
  std::vector<__nv_bfloat16 *> a_rm, a_t, p_rm, p_t;
  float *psi_out = nullptr;
  double *y = nullptr;
  long train_cap = 0, ld = 0;
  __nv_bfloat16 *act[2] = {nullptr, nullptr};
  double *d_out = nullptr;
  int8_t *d_codes = nullptr;
  long cap = 0;
  int max_width = 0;
  bool packed = false;
  cudaStream_t stream = nullptr;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int load_encode() {
  if (g_encode) return CAB_OK;
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr) { set_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e)); return CAB_ECUDA; }
  g_encode = (EncodeTiledFn) fn;
  return CAB_OK;
}

// row-major bf16 matrix [rows][cols] (cols contiguous, row stride ld elements; 0 = cols) -> tiled map with box
// {64 cols, box_rows}, 128-byte swizzle; reads past rows / cols are zero-filled
static int make_map(CUtensorMap *m, const void *ptr, long rows, long cols, int box_rows, long ld = 0) {
  const cuuint64_t gdim[2] = {(cuuint64_t) cols, (cuuint64_t) rows};
  const cuuint64_t gstride[1] = {(cuuint64_t) (ld ? ld : cols) * 2};
  const cuuint32_t box[2] = {(cuuint32_t) tc::BK, (cuuint32_t) box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%ld cols=%ld", (int) r, rows, cols); return CAB_ECUDA; }
  return CAB_OK;
}

static bool tc_supported(const cab_net *net) {
  const int L = (int) net->dims.size();
  if (L < 3 || net->dims[0] != 64 || net->dims[L - 1] != 1) return false;
  for (int l = 1; l + 1 < L; ++l)
    if (net->dims[l] % tc::BN != 0) return false;
  return true;
}

// ---- fused fp16 path for {64, N1, N2, N3, 1} nets (mlp_tc_fused.cuh) -------------------------------------------------
// A leaf-queue flush in launches of a few thousand boards is bound by the HOST's launch rate (a 4 096-board launch of the
// fused kernel runs ~5 us): cab_eval_f16_dev_batches captures its launch sequence once per (buffers, sizes, streams) into
// a CUDA graph and replays it.
constexpr int kGraphStreams = 8, kGraphCache = 8;
struct BatchGraph {
  const int8_t *codes = nullptr;
  double *out = nullptr;
  long n = 0, batch = 0;
  int n_streams = 0;
  cudaStream_t streams[kGraphStreams] = {};
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  uint64_t last_used = 0;
};
struct F16State {
  uint8_t *image = nullptr;
  tcf::PackArgs pack{};
  tcf::FusedArgs args{};
  size_t smem = 0;
  int8_t *d_codes = nullptr;
  double *d_out = nullptr;
  long cap = 0;
  cudaStream_t stream = nullptr;
  std::mutex graph_mu;
  std::vector<BatchGraph> graphs;
  uint64_t graph_clock = 0;
  cudaEvent_t ev_fork = nullptr, ev_join[kGraphStreams] = {}, ev_done = nullptr;
};

static bool f16_supported(const cab_net *net) {
  if (net->dims.size() != 5 || net->dims[0] != 64 || net->dims[4] != 1) return false;
  for (int l = 1; l <= 3; ++l)
    if (net->dims[l] < 1 || net->dims[l] > 256) return false;
  return true;
}

static void f16_delete(F16State *s) {
  if (!s) return;
  cudaFree(s->image); cudaFree(s->d_codes); cudaFree(s->d_out);
  if (s->stream) cudaStreamDestroy(s->stream);
  for (auto &g : s->graphs) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    if (g.graph) cudaGraphDestroy(g.graph);
  }
  if (s->ev_fork) cudaEventDestroy(s->ev_fork);
  if (s->ev_done) cudaEventDestroy(s->ev_done);
  for (auto e : s->ev_join) if (e) cudaEventDestroy(e);
  delete s;
}
static void f16_free(cab_net *net) {
  f16_delete(static_cast<F16State *>(net->f16.load()));
  net->f16 = nullptr;
}

// geometry of the image + device allocations of a new state
static int f16_construct(cab_net *net, F16State *s) {
  uint32_t off = 0;
  for (int l = 0; l < 3; ++l) {
    const int K = net->dims[l], N = net->dims[l + 1];
    const int kp = l == 0 ? 64 : (s->pack.n_pad[l - 1] + 63) / 64 * 64;   // next K = previous padded width, in 64-blocks
    const int np = (N + 15) / 16 * 16;
    s->pack.K[l] = K; s->pack.N[l] = N; s->pack.woff[l] = net->woff[l];
    s->pack.k_pad[l] = kp; s->pack.n_pad[l] = np; s->pack.w_off[l] = off;
    s->args.k_pad[l] = kp; s->args.n_pad[l] = np; s->args.w_off[l] = off;
    s->args.k_steps[l] = l == 0 ? 4 : s->pack.n_pad[l - 1] / 16;   // the A row holds exactly the previous padded width
    off += (uint32_t) np * (uint32_t) kp * 2u;     // np multiple of 8 and kp of 64: every K block is a whole number of 1 KB atoms
  }
  s->pack.K[3] = net->dims[3]; s->pack.N[3] = 1; s->pack.woff[3] = net->woff[3];
  s->pack.w_last_off = s->args.w_last_off = off;
  off += (uint32_t) s->pack.n_pad[2] * 4u;
  off = (off + 15u) & ~15u;
  s->args.image_bytes = off;
  s->smem = (size_t) ((off + 1023u) & ~1023u) + 2 * tcf::kBoardBytes + 1024 + 512 + 2 * 3 * tcf::kTile * sizeof(float);   // image, two board buffers, alignment slack, barriers + flags, partial dot products
  if (s->smem > 227 * 1024) { set_error("the fp16 weight image (%u bytes) does not fit shared memory", off); return CAB_EINVAL; }
  CAB_CUDA(cudaMalloc(&s->image, off));
  s->pack.image = s->image;
  s->args.image = s->image;
  CAB_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  CAB_CUDA(cudaFuncSetAttribute(tcf::mlp_tc_fused_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) s->smem));
  CAB_CUDA(cudaFuncSetAttribute(tcf::mlp_tc_fused_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) s->smem));
  return CAB_OK;
}

// Builds the state / the weight image on first use and after a weight change, under the net's lock (double-checked: the
// steady state takes no lock, so many host threads can launch concurrently).
static int f16_prepare(cab_net *net, cudaStream_t stream) {
  if (!f16_supported(net)) {
    set_error("the fused fp16 path needs a net shaped {64, N1, N2, N3, 1} with N1, N2, N3 <= 256");
    return CAB_EINVAL;
  }
  if (net->f16.load(std::memory_order_acquire) != nullptr && !net->f16_dirty) return CAB_OK;
  std::lock_guard<std::mutex> lk(net->mu);
  F16State *s = static_cast<F16State *>(net->f16.load());
  bool fresh = false;
  if (!s) {   // built aside and published only when complete: a failed construction leaves the net without a state
    s = new F16State();
    const int rc = f16_construct(net, s);
    if (rc != CAB_OK) { f16_delete(s); return rc; }
    fresh = true;
  }
  if (net->f16_dirty) {
    s->pack.w = net->d_w;
    tcf::f16_pack_kernel<<<64, 256, 0, stream ? stream : s->stream>>>(s->pack);
    net->launches++;
    const cudaError_t e1 = cudaGetLastError();
    const cudaError_t e2 = cudaStreamSynchronize(stream ? stream : s->stream);   // rare (weights changed); later launches may use other streams
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
      if (fresh) f16_delete(s);
      return cuda_fail(e1 != cudaSuccess ? e1 : e2, "f16_pack_kernel", __FILE__, __LINE__);
    }
    net->f16_dirty = false;
  }
  if (fresh) net->f16.store(s, std::memory_order_release);
  return CAB_OK;
}

// the launch itself; the caller has run f16_prepare
// max_ctas: the SMs this launch may count on (launches that run side by side on several streams share the device)
static int f16_launch(cab_net *net, const int8_t *d_codes, long n, double *d_out, cudaStream_t stream, int max_ctas = 0) {
  F16State *s = static_cast<F16State *>(net->f16.load(std::memory_order_acquire));
  tcf::FusedArgs a = s->args;
  a.codes = d_codes;
  a.n_boards = n;
  a.out = d_out;
  const long tiles = (n + tcf::kTile - 1) / tcf::kTile;
  static const bool profile = getenv("CAB_F16_PROFILE") != nullptr;   // debug: clock stamps of every CTA's second tile, printed to stderr
  long long *d_prof = nullptr;
  if (profile) {
    CAB_CUDA(cudaMalloc(&d_prof, (size_t) net->sm_count * 64 * sizeof(long long)));
    CAB_CUDA(memset_now(d_prof, 0, (size_t) net->sm_count * 64 * sizeof(long long)));
    a.prof = d_prof;
  }
  static const int ch = getenv("CAB_F16_CH") ? atoi(getenv("CAB_F16_CH")) : 16;   // accumulator columns per epilogue chunk (sweeps: 16 is best)
  // every CTA fetches the whole weight image and its first tile runs without overlap: with fewer SMs to count on, fewer
  // CTAs with more tiles each (equal shares)
  const long cap = max_ctas > 0 ? std::min(max_ctas, net->sm_count) : net->sm_count;
  const long tiles_per_cta = (tiles + cap - 1) / cap;
  const int grid = (int) ((tiles + tiles_per_cta - 1) / tiles_per_cta);
  cudaStream_t st = stream ? stream : s->stream;
  if (ch == 32) tcf::mlp_tc_fused_kernel<32><<<grid, tcf::kThreadsP, s->smem, st>>>(a);
  else tcf::mlp_tc_fused_kernel<16><<<grid, tcf::kThreadsP, s->smem, st>>>(a);
  net->launches++;
  CAB_CUDA(cudaGetLastError());
  if (profile) {
    CAB_CUDA(cudaDeviceSynchronize());
    std::vector<long long> h((size_t) net->sm_count * 64);
    CAB_CUDA(cudaMemcpy(h.data(), d_prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(d_prof);
    double e[8] = {}, m[10] = {}, ws[16] = {}, we[16] = {};
    int n_cta = 0;
    for (int c = 0; c < grid; ++c) {
      const long long *p = &h[(size_t) c * 64];
      if (p[0] == 0 || p[7] == 0 || p[16] == 0) continue;
      ++n_cta;
      for (int i = 0; i < 8; ++i) e[i] += (double) (p[i] - p[0]);
      for (int i = 0; i < 10; ++i) m[i] += (double) (p[16 + i] - p[0]);
      for (int i = 0; i < 16; ++i) { ws[i] += (double) (p[48 + i] - p[0]); we[i] += (double) (p[32 + i] - p[0]); }
    }
    if (n_cta > 0) {
      fprintf(stderr, "[f16 profile] boards=%ld ctas=%d  epilogue warp 0 (cycles after tile start): d0_full %.0f epi0_done %.0f d1_full %.0f epi1_done %.0f a0_stored %.0f d2_full %.0f final_done %.0f\n",
              n, n_cta, e[1] / n_cta, e[2] / n_cta, e[3] / n_cta, e[4] / n_cta, e[5] / n_cta, e[6] / n_cta, e[7] / n_cta);
      fprintf(stderr, "[f16 profile]   MMA thread: loop_top %.0f a0_seen %.0f L0_issued %.0f e0_first %.0f e0_last %.0f L1_committed %.0f e1_first %.0f e1_last %.0f L2_committed %.0f\n",
              m[0] / n_cta, m[1] / n_cta, m[2] / n_cta, m[3] / n_cta, m[4] / n_cta, m[5] / n_cta, m[6] / n_cta, m[7] / n_cta, m[8] / n_cta);
      fprintf(stderr, "[f16 profile]   MMA thread waited %.0f cycles on e_ready in the layer-2 loop; epilogue 1 start / end per warp:", (m[9] - m[0]) / n_cta);
      for (int i = 0; i < 16; ++i) fprintf(stderr, " %.0f/%.0f", ws[i] / n_cta, we[i] / n_cta);
      fprintf(stderr, "\n");
    }
  }
  return CAB_OK;
}
static int f16_forward(cab_net *net, const int8_t *d_codes, long n, double *d_out, cudaStream_t stream) {
  int rc = f16_prepare(net, stream);
  return rc != CAB_OK ? rc : f16_launch(net, d_codes, n, d_out, stream);
}

void tc_free(cab_net *net) {
  f16_free(net);
  if (net->tc_done) { cudaEventDestroy(net->tc_done); net->tc_done = nullptr; }
  TcState *s = static_cast<TcState *>(net->tc);
  if (!s) return;
  for (auto *p : s->wt) cudaFree(p);
  for (auto *p : s->wkn) cudaFree(p);
  for (auto *v : {&s->a_rm, &s->a_t, &s->p_rm, &s->p_t})
    for (auto *p : *v) cudaFree(p);
  cudaFree(s->psi_out); cudaFree(s->y);
  cudaFree(s->w_last); cudaFree(s->act[0]); cudaFree(s->act[1]); cudaFree(s->d_out); cudaFree(s->d_codes);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  net->tc = nullptr;
}

static int tc_prepare(cab_net *net, long n, cudaStream_t stream) {
  if (!tc_supported(net)) {
    set_error("bf16 tcgen05 path needs hidden widths that are multiples of %d (got a net it does not fit); use the fp64 path", tc::BN);
    return CAB_EINVAL;
  }
  int rc = load_encode();
  if (rc != CAB_OK) return rc;
  TcState *s = static_cast<TcState *>(net->tc);
  const int L = (int) net->dims.size();
  if (!s) {
    s = new TcState();
    net->tc = s;   // (under tc_mu; a failure below leaves a state tc_free releases and the next call completes: every
                   //  allocation further down is guarded by its own null / capacity test)
    for (int l = 0; l + 2 < L; ++l) {
      __nv_bfloat16 *p = nullptr;
      CAB_CUDA(cudaMalloc(&p, (size_t) net->dims[l] * net->dims[l + 1] * 2));
      s->wt.push_back(p);
      s->max_width = std::max(s->max_width, net->dims[l + 1]);
    }
    CAB_CUDA(cudaMalloc(&s->w_last, (size_t) net->dims[L - 2] * sizeof(float)));
    CAB_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CAB_CUDA(cudaFuncSetAttribute(tc::tc_gemm_bf16_kernel<tc::kEpiAct>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) tc::kSmemTc));
    CAB_CUDA(cudaFuncSetAttribute(tc::tc_gemm_bf16_kernel<tc::kEpiDeriv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) tc::kSmemTc));
    CAB_CUDA(cudaFuncSetAttribute(tc::tc_gemm_bf16_kernel<tc::kEpiGrad>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) tc::kSmemTc));
    CAB_CUDA(cudaFuncSetAttribute(tc::tc_gemm2_bf16_kernel<tc::kEpiAct>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) tc::kSmemTc2));
    CAB_CUDA(cudaFuncSetAttribute(tc::tc_gemm2_bf16_kernel<tc::kEpiDeriv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) tc::kSmemTc2));
    CAB_CUDA(cudaFuncSetAttribute(tc::tc_gemm2_bf16_kernel<tc::kEpiGrad>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) tc::kSmemTc2));
  }
  if (stream == nullptr) stream = s->stream;  // host entry point: everything on the state's own stream
  if (!s->packed || net->tc_dirty) {
    for (int l = 0; l + 2 < L; ++l) {
      const long total = (long) net->dims[l] * net->dims[l + 1];
      tc::pack_wt_bf16_kernel<<<(int) std::min<long>((total + 255) / 256, 148L * 16), 256, 0, stream>>>(
          net->d_w + net->woff[l], net->dims[l], net->dims[l + 1], s->wt[l]);
    }
    tc::pack_last_f32_kernel<<<8, 256, 0, stream>>>(net->d_w + net->woff[L - 2], net->dims[L - 2], s->w_last);
    net->launches += L - 1;
    CAB_CUDA(cudaGetLastError());
    CAB_CUDA(cudaStreamSynchronize(stream));  // rare (weights changed): later launches may sit on another stream
    s->packed = true;
    s->wkn_packed = false;
    net->tc_dirty = false;
  }
  if (n > s->cap) {
    cudaFree(s->act[0]); cudaFree(s->act[1]); cudaFree(s->d_out); cudaFree(s->d_codes);
    const long pad = (n + tc::BM - 1) / tc::BM * tc::BM;
    CAB_CUDA(cudaMalloc(&s->act[0], (size_t) pad * std::max(s->max_width, 64) * 2));
    CAB_CUDA(cudaMalloc(&s->act[1], (size_t) pad * std::max(s->max_width, 64) * 2));
    CAB_CUDA(cudaMalloc(&s->d_out, pad * sizeof(double)));
    CAB_CUDA(cudaMalloc(&s->d_codes, pad * 64));
    s->cap = pad;
  }
  return CAB_OK;
}

// CTA-pair kernels (cta_group::2, 256 x 256 tiles) pay off once there are several waves of pair-sized work items:
// measured forward 1.18 vs 1.14 PFLOP/s at 16 384 boards, 1.14 vs 1.08 at 65 536, but 0.84 vs 0.94 at 4 096.
// CAB_TC_PAIR=0 / 1 forces one kind.
static bool use_pair(long n_boards) {
  static const int forced = getenv("CAB_TC_PAIR") ? atoi(getenv("CAB_TC_PAIR")) : -1;
  return forced >= 0 ? forced != 0 : n_boards >= 12288;
}
static int b_box_rows(bool pair) { return pair ? tc::BN / 2 : tc::BN; }

static int launch_gemm(cab_net *net, int epi, const CUtensorMap &ma, const CUtensorMap &mb, const tc::TcGemmArgs &a, cudaStream_t stream,
                       bool pair) {
  if (pair) {
    const int items = ((a.M + 2 * tc::BM - 1) / (2 * tc::BM)) * (a.N / tc::BN) * std::max(1, a.split_k);
    const int grid = 2 * std::min(items, net->sm_count / 2);
    if (epi == tc::kEpiAct) tc::tc_gemm2_bf16_kernel<tc::kEpiAct><<<grid, tc::kThreadsTc, tc::kSmemTc2, stream>>>(ma, mb, a);
    else if (epi == tc::kEpiDeriv) tc::tc_gemm2_bf16_kernel<tc::kEpiDeriv><<<grid, tc::kThreadsTc, tc::kSmemTc2, stream>>>(ma, mb, a);
    else tc::tc_gemm2_bf16_kernel<tc::kEpiGrad><<<grid, tc::kThreadsTc, tc::kSmemTc2, stream>>>(ma, mb, a);
    net->launches++;
    CAB_CUDA(cudaGetLastError());
    return CAB_OK;
  }
  const int tiles = ((a.M + tc::BM - 1) / tc::BM) * (a.N / tc::BN);
  const int grid = std::min(tiles * std::max(1, a.split_k), net->sm_count);
  if (epi == tc::kEpiAct) tc::tc_gemm_bf16_kernel<tc::kEpiAct><<<grid, tc::kThreadsTc, tc::kSmemTc, stream>>>(ma, mb, a);
  else if (epi == tc::kEpiDeriv) tc::tc_gemm_bf16_kernel<tc::kEpiDeriv><<<grid, tc::kThreadsTc, tc::kSmemTc, stream>>>(ma, mb, a);
  else tc::tc_gemm_bf16_kernel<tc::kEpiGrad><<<grid, tc::kThreadsTc, tc::kSmemTc, stream>>>(ma, mb, a);
  net->launches++;
  CAB_CUDA(cudaGetLastError());
  return CAB_OK;
}

// save = false: ping-pong buffers (evaluation). save = true: every layer's activation is kept row-major (a_rm) and,
// where the dW GEMM will contract over boards, transposed (a_t) — propagate_for_training, network.cpp:139-153.
static int tc_forward(cab_net *net, const int8_t *d_codes, long n, double *d_out, cudaStream_t stream, bool save = false) {
  int rc = tc_prepare(net, n, stream);
  if (rc != CAB_OK) return rc;
  TcState *s = static_cast<TcState *>(net->tc);
  const int L = (int) net->dims.size();
  __nv_bfloat16 *x0 = save ? s->a_rm[0] : s->act[0];
  tc::codes_to_bf16_kernel<<<(int) std::min<long>((n * 64 + 255) / 256, 148L * 16), 256, 0, stream>>>(d_codes, n * 64, x0);
  net->launches++;
  if (save) {
    tc::codes_to_bf16_t_kernel<<<(int) std::min<long>((n * 64 + 255) / 256, 148L * 16), 256, 0, stream>>>(d_codes, n, s->ld, s->a_t[0]);
    net->launches++;
  }
  int cur = 0;
  const bool pair = use_pair(n);
  const __nv_bfloat16 *in = x0;
  for (int l = 0; l + 2 < L; ++l) {
    const int K = net->dims[l], N = net->dims[l + 1];
    CUtensorMap mx, mw;
    if ((rc = make_map(&mx, in, n, K, tc::BM)) != CAB_OK) return rc;
    if ((rc = make_map(&mw, s->wt[l], N, K, b_box_rows(pair))) != CAB_OK) return rc;
    tc::TcGemmArgs a{};
    a.M = (int) n; a.N = N; a.K = K;
    a.Y = save ? s->a_rm[l + 1] : s->act[1 - cur];
    a.Yt = (save && l + 3 < L) ? s->a_t[l + 1] : nullptr;     // a_{L-2} feeds only the last layer's GEMV
    a.ldt = s->ld;
    a.split_k = 1;
    a.apply_act = 1;
    if ((rc = launch_gemm(net, tc::kEpiAct, mx, mw, a, stream, pair)) != CAB_OK) return rc;
    in = a.Y;
    cur = 1 - cur;
  }
  const int K = net->dims[L - 2];
  tc::gemv_out_kernel<<<(int) ((n * 32 + 255) / 256), 256, 0, stream>>>(in, s->w_last, n, K, d_out);
  net->launches++;
  CAB_CUDA(cudaGetLastError());
  return CAB_OK;
}

static int tc_train_prepare(cab_net *net, long n, cudaStream_t stream) {
  int rc = tc_prepare(net, n, stream);
  if (rc != CAB_OK) return rc;
  TcState *s = static_cast<TcState *>(net->tc);
  const int L = (int) net->dims.size();
  if (s->wkn.empty()) {
    s->wkn.assign(L - 2, nullptr);
    for (int l = 1; l + 2 < L; ++l) CAB_CUDA(cudaMalloc(&s->wkn[l], (size_t) net->dims[l] * net->dims[l + 1] * 2));
  }
  if (!s->wkn_packed) {
    for (int l = 1; l + 2 < L; ++l) {
      const long total = (long) net->dims[l] * net->dims[l + 1];
      tc::pack_w_bf16_kernel<<<(int) std::min<long>((total + 255) / 256, 148L * 16), 256, 0, stream>>>(net->d_w + net->woff[l], total, s->wkn[l]);
      net->launches++;
    }
    CAB_CUDA(cudaGetLastError());
    s->wkn_packed = true;
  }
  if (n > s->train_cap) {
    for (auto *v : {&s->a_rm, &s->a_t, &s->p_rm, &s->p_t}) {
      for (auto *p : *v) cudaFree(p);
      v->assign(L - 1, nullptr);
    }
    cudaFree(s->psi_out); cudaFree(s->y);
    const long ld = (n + 2 * tc::BM - 1) / (2 * tc::BM) * (2 * tc::BM);   // whole 256-row tiles: pad rows of the transposed copies stay inside a row
    for (int l = 0; l + 1 < L; ++l) {
      const size_t bytes = (size_t) ld * net->dims[l] * 2;
      CAB_CUDA(cudaMalloc(&s->a_rm[l], bytes));
      if (l + 2 < L) { CAB_CUDA(cudaMalloc(&s->a_t[l], bytes)); CAB_CUDA(memset_now(s->a_t[l], 0, bytes)); }
      if (l >= 1) {
        CAB_CUDA(cudaMalloc(&s->p_rm[l], bytes));
        CAB_CUDA(cudaMalloc(&s->p_t[l], bytes));
        CAB_CUDA(memset_now(s->p_t[l], 0, bytes));
      }
    }
    CAB_CUDA(cudaMalloc(&s->psi_out, ld * sizeof(float)));
    CAB_CUDA(cudaMalloc(&s->y, ld * sizeof(double)));
    CAB_CUDA(cudaDeviceSynchronize());   // the memsets ran on the legacy stream
    s->train_cap = n;
    s->ld = ld;
  }
  return CAB_OK;
}

// Forward with saved activations + backward, leaving sum_b dW in grad[0 .. nw) (fp64, the reference's n,i,j order) and
// sum_b E in grad[nw]. The caller zeroed grad. Mirrors K2 -> mlp_backward_kernel -> grad_kernel of the fp64 path.
int tc_grad(cab_net *net, const int8_t *d_codes, const double *d_targets, long n, double *d_grad, cudaStream_t stream) {
  int rc = tc_train_prepare(net, n, stream);
  if (rc != CAB_OK) return rc;
  TcState *s = static_cast<TcState *>(net->tc);
  const int L = (int) net->dims.size();
  if ((rc = tc_forward(net, d_codes, n, s->y, stream, /*save=*/true)) != CAB_OK) return rc;
  // output neuron and the last (N = 1) layer
  const int KL = net->dims[L - 2];
  tc::out_delta_kernel<<<(int) std::min<long>((n + 255) / 256, 148L * 4), 256, 0, stream>>>(s->y, d_targets, n, s->psi_out, d_grad + net->n_weights);
  const long slice = std::max<long>(64, (n + 63) / 64);
  tc::gemv_t_kernel<<<dim3((KL / 2 + 255) / 256, (unsigned) ((n + slice - 1) / slice)), 256, 0, stream>>>(
      s->a_rm[L - 2], s->psi_out, n, KL, slice, d_grad + net->woff[L - 2]);
  tc::psi_last_kernel<<<dim3((unsigned) ((n + 63) / 64), KL / 64), 256, 0, stream>>>(s->a_rm[L - 2], s->w_last, s->psi_out, n, KL,
                                                                                    s->p_rm[L - 2], s->p_t[L - 2], s->ld);
  net->launches += 3;
  CAB_CUDA(cudaGetLastError());
  const bool pair = use_pair(n);
  for (int l = L - 3; l >= 0; --l) {
    const int K = net->dims[l], N = net->dims[l + 1];
    {  // dW_l[K][N] = a_l^T psi_{l+1}: contraction over boards, both operands read from their transposed copies
      CUtensorMap ma, mb;
      if ((rc = make_map(&ma, s->a_t[l], K, n, tc::BM, s->ld)) != CAB_OK) return rc;
      if ((rc = make_map(&mb, s->p_t[l + 1], N, n, b_box_rows(pair), s->ld)) != CAB_OK) return rc;
      tc::TcGemmArgs a{};
      a.M = K; a.N = N; a.K = (int) n;
      a.G = d_grad + net->woff[l];
      const int tile_rows = pair ? 2 * tc::BM : tc::BM;
      const int tiles = ((K + tile_rows - 1) / tile_rows) * (N / tc::BN);
      const int kblocks = (int) ((n + tc::BK - 1) / tc::BK);
      // narrow layers (the 64-wide input) have too few output tiles to fill the chip: split the board dimension
      const int slots = pair ? net->sm_count / 2 : net->sm_count;   // concurrently running work items
      a.split_k = tiles * 2 > slots ? 1 : std::max(1, std::min(kblocks / 8, slots / tiles));
      if ((rc = launch_gemm(net, tc::kEpiGrad, ma, mb, a, stream, pair)) != CAB_OK) return rc;
    }
    if (l >= 1) {  // psi_l = (psi_{l+1} W_l^T) * f'(a_l); the input layer needs no psi (network.cpp:248: L stops at 0)
      CUtensorMap ma, mb;
      if ((rc = make_map(&ma, s->p_rm[l + 1], n, N, tc::BM)) != CAB_OK) return rc;
      if ((rc = make_map(&mb, s->wkn[l], K, N, b_box_rows(pair))) != CAB_OK) return rc;
      tc::TcGemmArgs a{};
      a.M = (int) n; a.N = K; a.K = N;
      a.Y = s->p_rm[l]; a.Yt = s->p_t[l]; a.ldt = s->ld;
      a.aux = s->a_rm[l];
      a.split_k = 1;
      if ((rc = launch_gemm(net, tc::kEpiDeriv, ma, mb, a, stream, pair)) != CAB_OK) return rc;
    }
  }
  return CAB_OK;
}

// The bf16 path works in per-net scratch (packed weights, ping-pong activations, staging), so its calls take turns:
// the caller holds net->tc_mu from tc_turn_begin to tc_turn_end, and the device work of a call is ordered after the
// previous call's on whatever stream that one ran (event), so concurrent host threads / streams stay correct.
int tc_turn_begin(cab_net *net, cudaStream_t stream) {
  if (net->tc_done == nullptr) { CAB_CUDA(cudaEventCreateWithFlags(&net->tc_done, cudaEventDisableTiming)); return CAB_OK; }
  CAB_CUDA(cudaStreamWaitEvent(stream, net->tc_done, 0));
  return CAB_OK;
}
int tc_turn_end(cab_net *net, cudaStream_t stream) {
  CAB_CUDA(cudaEventRecord(net->tc_done, stream));
  return CAB_OK;
}

int tc_forward_out(cab_net *net, const int8_t *d_codes, long n, double *d_out, cudaStream_t stream) {
  return tc_forward(net, d_codes, n, d_out, stream);
}

bool tc_fits(const cab_net *net) { return tc_supported(net); }

}  // namespace cab

using namespace cab;

extern "C" {

int cab_eval_bf16_dev(cab_net *net, const int8_t *codes_dev, long n, double *out_dev, void *stream) {
  if (!net || n < 0 || (n > 0 && (!codes_dev || !out_dev))) { set_error("bad argument"); return CAB_EINVAL; }
  if (n == 0) return CAB_OK;
  CAB_CUDA(cudaSetDevice(net->device));
  std::lock_guard<std::mutex> turn(net->tc_mu);
  int rc = tc_prepare(net, n, (cudaStream_t) stream);   // creates the state's own stream, used when `stream` is null
  if (rc != CAB_OK) return rc;
  cudaStream_t s = stream ? (cudaStream_t) stream : static_cast<TcState *>(net->tc)->stream;
  if ((rc = tc_turn_begin(net, s)) != CAB_OK) return rc;
  if ((rc = tc_forward(net, codes_dev, n, out_dev, s)) != CAB_OK) return rc;
  return tc_turn_end(net, s);
}

int cab_eval_f16_dev(cab_net *net, const int8_t *codes_dev, long n, double *out_dev, void *stream) {
  if (!net || n < 0 || (n > 0 && (!codes_dev || !out_dev))) { set_error("bad argument"); return CAB_EINVAL; }
  if (n == 0) return CAB_OK;
  CAB_CUDA(cudaSetDevice(net->device));
  return f16_forward(net, codes_dev, n, out_dev, (cudaStream_t) stream);
}

// The launches of one cab_eval_f16_dev_batches call as a CUDA graph: captured on the caller's streams (fork from
// streams[0], round-robin launches, join), cached by (buffers, sizes, streams), replayed on streams[0]; the other
// streams then wait for the replay, so every caller stream still orders behind the whole call.
static int f16_batches_graph(cab_net *net, F16State *s, const int8_t *codes_dev, long n, long batch, double *out_dev,
                             void *const *streams, int n_streams) {
  std::lock_guard<std::mutex> lk(s->graph_mu);
  if (s->ev_fork == nullptr) {
    CAB_CUDA(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
    CAB_CUDA(cudaEventCreateWithFlags(&s->ev_done, cudaEventDisableTiming));
    for (auto &e : s->ev_join) CAB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  BatchGraph *g = nullptr;
  for (auto &c : s->graphs) {
    bool same = c.codes == codes_dev && c.out == out_dev && c.n == n && c.batch == batch && c.n_streams == n_streams;
    for (int i = 0; same && i < n_streams; ++i) same = c.streams[i] == (cudaStream_t) streams[i];
    if (same) { g = &c; break; }
  }
  const long n_launches = (n + batch - 1) / batch;
  if (g == nullptr) {
    if ((int) s->graphs.size() >= kGraphCache) {            // evict the least recently used
      size_t old = 0;
      for (size_t i = 1; i < s->graphs.size(); ++i) if (s->graphs[i].last_used < s->graphs[old].last_used) old = i;
      cudaGraphExecDestroy(s->graphs[old].exec);
      cudaGraphDestroy(s->graphs[old].graph);
      s->graphs.erase(s->graphs.begin() + old);
    }
    BatchGraph c;
    c.codes = codes_dev; c.out = out_dev; c.n = n; c.batch = batch; c.n_streams = n_streams;
    for (int i = 0; i < n_streams; ++i) c.streams[i] = (cudaStream_t) streams[i];
    cudaStream_t s0 = c.streams[0];
    CAB_CUDA(cudaStreamBeginCapture(s0, cudaStreamCaptureModeThreadLocal));
    int rc = CAB_OK;
    cudaError_t e = cudaEventRecord(s->ev_fork, s0);
    for (int i = 1; e == cudaSuccess && i < n_streams; ++i) e = cudaStreamWaitEvent(c.streams[i], s->ev_fork, 0);
    long k = 0;
    const int share = std::max(1, net->sm_count / (int) std::min<long>(n_streams, n_launches));   // SMs per concurrent launch
    for (long b0 = 0; e == cudaSuccess && rc == CAB_OK && b0 < n; b0 += batch, ++k)
      rc = f16_launch(net, codes_dev + b0 * 64, std::min(batch, n - b0), out_dev + b0, c.streams[k % n_streams], share);
    for (int i = 1; e == cudaSuccess && i < n_streams; ++i) {
      e = cudaEventRecord(s->ev_join[i], c.streams[i]);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(s0, s->ev_join[i], 0);
    }
    const cudaError_t e_end = cudaStreamEndCapture(s0, &c.graph);
    if (rc != CAB_OK) { if (c.graph) cudaGraphDestroy(c.graph); return rc; }
    if (e != cudaSuccess || e_end != cudaSuccess) { if (c.graph) cudaGraphDestroy(c.graph); return cuda_fail(e != cudaSuccess ? e : e_end, "graph capture", __FILE__, __LINE__); }
    net->launches -= n_launches;                             // f16_launch counted the captured launches; the replay below counts them
    e = cudaGraphInstantiate(&c.exec, c.graph, 0);
    if (e != cudaSuccess) { cudaGraphDestroy(c.graph); return cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__); }
    s->graphs.push_back(c);
    g = &s->graphs.back();
  }
  g->last_used = ++s->graph_clock;
  CAB_CUDA(cudaGraphLaunch(g->exec, g->streams[0]));
  net->launches += n_launches;
  if (n_streams > 1) {
    CAB_CUDA(cudaEventRecord(s->ev_done, g->streams[0]));
    for (int i = 1; i < n_streams; ++i) CAB_CUDA(cudaStreamWaitEvent(g->streams[i], s->ev_done, 0));
  }
  return CAB_OK;
}

int cab_eval_f16_dev_batches(cab_net *net, const int8_t *codes_dev, long n, long batch, double *out_dev, void *const *streams,
                             int n_streams) {
  if (!net || n < 0 || batch <= 0 || n_streams <= 0 || !streams || (n > 0 && (!codes_dev || !out_dev))) { set_error("bad argument"); return CAB_EINVAL; }
  CAB_CUDA(cudaSetDevice(net->device));
  int rc = f16_prepare(net, (cudaStream_t) streams[0]);     // weights packed (and that stream drained) before any launch
  if (rc != CAB_OK) return rc;
  // many small launches: replay them as one graph (real streams only: the legacy default stream cannot be captured)
  static const bool use_graph = getenv("CAB_F16_NO_GRAPH") == nullptr && getenv("CAB_F16_PROFILE") == nullptr;
  bool real_streams = n_streams <= kGraphStreams;
  for (int i = 0; real_streams && i < n_streams; ++i) real_streams = streams[i] != nullptr;
  if (use_graph && real_streams && (n + batch - 1) / batch >= 8) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing((cudaStream_t) streams[0], &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone)
      return f16_batches_graph(net, static_cast<F16State *>(net->f16.load(std::memory_order_acquire)), codes_dev, n, batch, out_dev, streams, n_streams);
  }
  int i = 0;
  for (long b0 = 0; b0 < n; b0 += batch, ++i) {
    rc = f16_forward(net, codes_dev + b0 * 64, std::min(batch, n - b0), out_dev + b0, (cudaStream_t) streams[i % n_streams]);
    if (rc != CAB_OK) return rc;
  }
  return CAB_OK;
}

int cab_eval_f16_host(cab_net *net, const int8_t *codes, long n, double *out) {
  if (!net || n < 0 || (n > 0 && (!codes || !out))) { set_error("bad argument"); return CAB_EINVAL; }
  if (n == 0) return CAB_OK;
  CAB_CUDA(cudaSetDevice(net->device));
  int rc = f16_prepare(net, nullptr);
  if (rc != CAB_OK) return rc;
  F16State *s = static_cast<F16State *>(net->f16.load(std::memory_order_acquire));
  static std::mutex staging_mu;                  // the staging buffers are per net; host callers take turns
  std::lock_guard<std::mutex> lk(staging_mu);
  if (n > s->cap) {
    cudaFree(s->d_codes); cudaFree(s->d_out);
    s->d_codes = nullptr; s->d_out = nullptr;
    CAB_CUDA(cudaMalloc(&s->d_codes, n * 64));
    CAB_CUDA(cudaMalloc(&s->d_out, n * sizeof(double)));
    s->cap = n;
  }
  CAB_CUDA(cudaMemcpyAsync(s->d_codes, codes, n * 64, cudaMemcpyHostToDevice, s->stream));
  rc = f16_forward(net, s->d_codes, n, s->d_out, s->stream);
  if (rc != CAB_OK) return rc;
  CAB_CUDA(cudaMemcpyAsync(out, s->d_out, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CAB_CUDA(cudaStreamSynchronize(s->stream));
  return CAB_OK;
}

int cab_eval_bf16_host(cab_net *net, const int8_t *codes, long n, double *out) {
  if (!net || n < 0 || (n > 0 && (!codes || !out))) { set_error("bad argument"); return CAB_EINVAL; }
  if (n == 0) return CAB_OK;
  CAB_CUDA(cudaSetDevice(net->device));
  std::lock_guard<std::mutex> turn(net->tc_mu);
  int rc = tc_prepare(net, n, nullptr);
  if (rc != CAB_OK) return rc;
  TcState *s = static_cast<TcState *>(net->tc);
  if ((rc = tc_turn_begin(net, s->stream)) != CAB_OK) return rc;
  CAB_CUDA(cudaMemcpyAsync(s->d_codes, codes, n * 64, cudaMemcpyHostToDevice, s->stream));
  rc = tc_forward(net, s->d_codes, n, s->d_out, s->stream);
  if (rc != CAB_OK) return rc;
  CAB_CUDA(cudaMemcpyAsync(out, s->d_out, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if ((rc = tc_turn_end(net, s->stream)) != CAB_OK) return rc;
  CAB_CUDA(cudaStreamSynchronize(s->stream));
  return CAB_OK;
}

}  // extern "C"
