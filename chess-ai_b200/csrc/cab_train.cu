// cab_train.cu — training entry points of include/chessai_b200.h (Optimizer::optimize / dropout, train loop, DP).
#include <dlfcn.h>

#include <cstdio>
#include <algorithm>
#include <atomic>
#include <cstring>

#include "cab_internal.cuh"
#include "train_kernels.cuh"

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


// ---- NCCL through dlopen: the library must load (and evaluate) on machines without NCCL -------------------------
struct Id128 { char b[128]; };  // ncclUniqueId: 128 opaque bytes, passed BY VALUE to ncclCommInitRank
struct NcclApi {
  void *h = nullptr;
  int (*GetUniqueId)(void *) = nullptr;
  int (*CommInitRank)(void **, int, Id128, int) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.h != nullptr) return CAB_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    g_nccl.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  if (!g_nccl.h) { set_error("NCCL not found (dlopen libnccl.so.2): %s", dlerror()); return CAB_ENCCL; }
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId)) dlsym(g_nccl.h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank)) dlsym(g_nccl.h, "ncclCommInitRank");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce)) dlsym(g_nccl.h, "ncclAllReduce");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy)) dlsym(g_nccl.h, "ncclCommDestroy");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString)) dlsym(g_nccl.h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
    set_error("NCCL library lacks required symbols");
    return CAB_ENCCL;
  }
  return CAB_OK;
}
static int nccl_fail(int rc, const char *what) {
  set_error("NCCL error %d in %s: %s", rc, what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  return CAB_ENCCL;
}
constexpr int kNcclDouble = 8, kNcclSum = 0;  // ncclFloat64, ncclSum (nccl.h enums, stable across 2.x)

// ---- training buffers ------------------------------------------------------------------------------------------
static int ensure_train_state(cab_net *net) {
  if (net->train_stream == nullptr) CAB_CUDA(cudaStreamCreateWithFlags(&net->train_stream, cudaStreamNonBlocking));
  if (net->d_dw == nullptr) {
    CAB_CUDA(cudaMalloc(&net->d_dw, net->n_weights * sizeof(double)));
    CAB_CUDA(memset_now(net->d_dw, 0, net->n_weights * sizeof(double)));
  }
  if (net->d_grad == nullptr) {
    CAB_CUDA(cudaMalloc(&net->d_grad, (net->n_weights + 4) * sizeof(double)));
    CAB_CUDA(memset_now(net->d_grad, 0, (net->n_weights + 4) * sizeof(double)));
  }
  if (net->d_scalars == nullptr) {
    CAB_CUDA(cudaMalloc(&net->d_scalars, 64 * sizeof(double)));
    CAB_CUDA(memset_now(net->d_scalars, 0, 64 * sizeof(double)));
  }
  if (net->d_ctl == nullptr) {
    CAB_CUDA(cudaMalloc(&net->d_ctl, sizeof(TrainCtl)));
    CAB_CUDA(memset_now(net->d_ctl, 0, sizeof(TrainCtl)));
  }
  return CAB_OK;
}

static int ensure_batch_buffers(cab_net *net, long n) {
  const long pad = (n + 63) / 64 * 64;
  if (pad > net->train_cap) {
    cudaFree(net->d_acts); cudaFree(net->d_psis); cudaFree(net->d_tout);
    net->d_acts = net->d_psis = net->d_tout = nullptr;
    const size_t bytes = (size_t) net->act_units_total * pad * 8 * sizeof(double);
    CAB_CUDA(cudaMalloc(&net->d_acts, bytes));
    CAB_CUDA(cudaMalloc(&net->d_psis, bytes));
    CAB_CUDA(cudaMalloc(&net->d_tout, pad * sizeof(double)));
    CAB_CUDA(memset_now(net->d_acts, 0, bytes));
    CAB_CUDA(memset_now(net->d_psis, 0, bytes));
    net->train_cap = pad;
  }
  return CAB_OK;
}

static int ensure_host_staging(cab_net *net, long n) {
  if (n > net->tbuf_cap) {
    cudaFree(net->d_tcodes); cudaFree(net->d_ttargets);
    CAB_CUDA(cudaMalloc(&net->d_tcodes, n * 64));
    CAB_CUDA(cudaMalloc(&net->d_ttargets, n * sizeof(double)));
    net->tbuf_cap = n;
  }
  return CAB_OK;
}

// grad tasks (layer x tile blocks), built once per net and cached on the device
struct GradPlan {
  GradTask *d_tasks = nullptr;
  int n_tasks = 0;
  std::vector<GradTask> tasks;
  GradItem *d_items = nullptr;   // work items for `items_n` boards (rebuilt when the batch size changes)
  int n_items = 0;
  long items_n = -1;
};
static std::mutex g_plan_mu;
static std::vector<std::pair<cab_net *, GradPlan>> g_grad_plans;

// Splits the n boards (rounded up to 4) into slices and every slice into one item per task (a warp's unit of work).
// Items are ordered SLICE-MAJOR: the persistent warps draw them in order, so the ~1 200 items in flight at any moment
// cover the same few dozen board slices and every saved-activation / psi unit a slice needs is read from DRAM once and
// from L2 by the other tile blocks that contract over it (task-major order, round 2's first form, re-read the units
// once per tile block that needs them: 1.1 GB per 65 536 boards against 0.53 GB of operands, L2 hit 50 %). Inside a
// slice the expensive blocks come first, so the last items drawn are the thin ones. Slices start on stage boundaries
// (multiples of kGradStageBoards) and shrink with the batch so that a small batch still yields ~4 items per warp.
static int build_grad_items(cab_net *net, GradPlan &gp, long n) {
  if (gp.items_n == n) return CAB_OK;
  const long n4 = (n + 3) / 4 * 4;
  static const long env_slice = getenv("CAB_GRAD_SLICE") ? atol(getenv("CAB_GRAD_SLICE")) : 256;
  static const bool task_major = getenv("CAB_GRAD_TASK_MAJOR") != nullptr;   // the old order, kept for A/B timing
  const double target_items = (double) net->sm_count * kGradWarps * 4;
  const long slices_wanted = std::max<long>(1, std::lround(target_items / std::max(1, gp.n_tasks)));
  long per = std::min<long>(std::max<long>(64, env_slice), std::max<long>(64, (n4 + slices_wanted - 1) / slices_wanted));
  per = (per + kGradStageBoards - 1) / kGradStageBoards * kGradStageBoards;
  std::vector<int> order(gp.n_tasks);
  for (int i = 0; i < gp.n_tasks; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
    return gp.tasks[x].nkt * gp.tasks[x].nnt > gp.tasks[y].nkt * gp.tasks[y].nnt;
  });
  std::vector<GradItem> items;
  if (!task_major) {
    for (long b = 0; b < n4; b += per)
      for (int ti : order) items.push_back(GradItem{ti, (int) b, (int) std::min(b + per, n4), 0});
  } else {
    for (int ti = 0; ti < gp.n_tasks; ++ti)
      for (long b = 0; b < n4; b += per) items.push_back(GradItem{ti, (int) b, (int) std::min(b + per, n4), 0});
  }
  cudaFree(gp.d_items);
  gp.d_items = nullptr;
  CAB_CUDA(cudaMalloc(&gp.d_items, items.size() * sizeof(GradItem)));
  CAB_CUDA(upload_now(gp.d_items, items.data(), items.size() * sizeof(GradItem)));
  gp.n_items = (int) items.size();
  gp.items_n = n;
  return CAB_OK;
}

static int get_grad_plan(cab_net *net, long n, GradPlan *out) {
  std::lock_guard<std::mutex> lk(g_plan_mu);
  for (auto &p : g_grad_plans)
    if (p.first == net) {
      int rc = build_grad_items(net, p.second, n);
      *out = p.second;
      return rc;
    }
  std::vector<GradTask> tasks;
  const int L = (int) net->dims.size();
  for (int l = 0; l + 1 < L; ++l) {
    const int K = net->dims[l], N = net->dims[l + 1];
    const int kt = (K + 7) / 8, nt = (N + 7) / 8;
    // even split of the tile grid into blocks of at most kGK x kGN tiles (18 -> 4,4,4,3,3 rather than 4,4,4,4,2)
    const int gk = (kt + kGK - 1) / kGK, gn = (nt + kGN - 1) / kGN;
    for (int bk = 0; bk < gk; ++bk)
      for (int bn = 0; bn < gn; ++bn) {
        const int k0 = kt * bk / gk, k1 = kt * (bk + 1) / gk, n0 = nt * bn / gn, n1 = nt * (bn + 1) / gn;
        GradTask t{};
        t.a_unit = net->act_unit_off[l] + k0;
        t.p_unit = net->act_unit_off[l + 1] + n0;
        t.nkt = k1 - k0;
        t.nnt = n1 - n0;
        t.k0 = k0 * 8; t.n0 = n0 * 8;
        t.K = K; t.N = N;
        t.w_off = net->woff[l];
        tasks.push_back(t);
      }
  }
  GradPlan gp;
  gp.n_tasks = (int) tasks.size();
  CAB_CUDA(cudaMalloc(&gp.d_tasks, tasks.size() * sizeof(GradTask)));
  CAB_CUDA(upload_now(gp.d_tasks, tasks.data(), tasks.size() * sizeof(GradTask)));
  gp.tasks = tasks;
  int rc = build_grad_items(net, gp, n);
  g_grad_plans.push_back({net, gp});
  *out = gp;
  return rc;
}

void drop_grad_plan(cab_net *net) {
  std::lock_guard<std::mutex> lk(g_plan_mu);
  for (size_t i = 0; i < g_grad_plans.size(); ++i)
    if (g_grad_plans[i].first == net) {
      cudaFree(g_grad_plans[i].second.d_tasks);
      cudaFree(g_grad_plans[i].second.d_items);
      g_grad_plans.erase(g_grad_plans.begin() + i);
      return;
    }
}

static int launch_backward(cab_net *net, const double *d_targets, long b0, long n, cudaStream_t stream) {
  BackwardArgs a{};
  a.pack = reinterpret_cast<const uint8_t *>(net->bwd.d_pack);
  a.chunks = net->bwd.d_chunks;
  a.n_chunks = (int) net->bwd.chunks.size();
  a.a_units = net->bwd.a_units;
  a.n_boards = n;
  const Geometry geo = choose_geometry(a.a_units, net->stage_bytes, 0, net->sm_count);
  a.n_stages = geo.n_stages;
  a.stage_bytes = net->stage_bytes;
  const int boards_per_cta = geo.n_cons * kBoardsPerWarp;
  a.n_groups = (n + boards_per_cta - 1) / boards_per_cta;
  a.targets = d_targets + b0;
  a.acts = net->d_acts + b0 * 8;      // unit layout [unit][board][8]: a board offset is 8 doubles in every unit row
  a.psis = net->d_psis + b0 * 8;
  a.save_stride = net->train_cap * 8;
  a.out_unit = net->act_unit_off.back();
  a.err_sum = net->d_grad + net->n_weights;  // [sumE, sumE', count, spare]
  a.blob = net->bwd.d_blob;
  for (int st = 0; st < geo.n_stages && st < kMaxStages; ++st) {   // the ring's first fills, as kernel parameters
    const ChunkDesc &cd = net->bwd.chunks[st % net->bwd.chunks.size()];
    a.fill_off[st] = cd.src_off;
    a.fill_bytes[st] = cd.bytes;
  }
  const size_t smem = forward_smem_bytes(a.a_units, geo.n_cons, geo.n_stages, net->stage_bytes);
  const int grid = (int) std::min<long>(a.n_groups, (long) net->sm_count * geo.ctas_per_sm);
  static std::atomic<uint64_t> attr_devices{0};   // the attribute is per device: one bit per device ordinal
  if (!(attr_devices.load() >> (net->device & 63) & 1)) {
    CAB_CUDA(cudaFuncSetAttribute(mlp_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_devices.fetch_or(1ull << (net->device & 63));
  }
  mlp_backward_kernel<<<grid, geo.n_cons * 64, smem, stream>>>(a);
  net->launches++;
  CAB_CUDA(cudaGetLastError());
  return CAB_OK;
}

static int launch_grad(cab_net *net, long n, cudaStream_t stream) {
  GradPlan gp;
  int rc = get_grad_plan(net, n, &gp);
  if (rc != CAB_OK) return rc;
  GradArgs a{};
  a.tasks = gp.d_tasks;
  a.items = gp.d_items;
  a.n_items = gp.n_items;
  static const int env_warps = getenv("CAB_GRAD_WARPS") ? atoi(getenv("CAB_GRAD_WARPS")) : kGradWarps;
  static const int env_stages = getenv("CAB_GRAD_STAGES") ? atoi(getenv("CAB_GRAD_STAGES")) : kGradStages;
  a.n_warps = std::max(1, std::min(env_warps, kGradWarps));
  a.n_stages = std::max(2, env_stages);
  const size_t smem = (size_t) a.n_warps * a.n_stages * kGradStageBytes + (size_t) a.n_warps * 8 * sizeof(uint64_t);
  if (smem > 227 * 1024) { set_error("grad kernel geometry needs %zu bytes of shared memory", smem); return CAB_EINVAL; }
  a.acts = net->d_acts;
  a.psis = net->d_psis;
  a.save_stride = net->train_cap * 8;
  a.grad = net->d_grad;
  static std::atomic<uint64_t> attr_devices{0};
  if (!(attr_devices.load() >> (net->device & 63) & 1)) {
    CAB_CUDA(cudaFuncSetAttribute(grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_devices.fetch_or(1ull << (net->device & 63));
  }
  a.next_item = reinterpret_cast<int *>(net->d_scalars + 16);
  CAB_CUDA(cudaMemsetAsync(a.next_item, 0, sizeof(int), stream));
  const int grid = std::min((a.n_items + a.n_warps - 1) / a.n_warps, net->sm_count);
  grad_kernel<<<grid, a.n_warps * 32, smem, stream>>>(a);
  net->launches++;
  CAB_CUDA(cudaGetLastError());
  return CAB_OK;
}

// Big passes are issued as pieces of kPieceBoards boards round-robin over helper streams (event fork / join on the
// caller's stream): one persistent launch makes all 148 CTAs request the same weight chunk at the same moment and
// waits on L2 (measured 1.62e8 boards/s), desynchronised pieces do not (1.80e8).
constexpr long kPieceBoards = 8192;
template <typename F>
static int fan_out(cab_net *net, cudaStream_t stream, long n, F &&piece) {
  if (n <= 2 * kPieceBoards) return piece(0, n, stream);
  for (int i = 0; i < 4; ++i)
    if (net->aux_stream[i] == nullptr) {
      CAB_CUDA(cudaStreamCreateWithFlags(&net->aux_stream[i], cudaStreamNonBlocking));
      CAB_CUDA(cudaEventCreateWithFlags(&net->aux_event[i], cudaEventDisableTiming));
    }
  if (net->aux_event[4] == nullptr) CAB_CUDA(cudaEventCreateWithFlags(&net->aux_event[4], cudaEventDisableTiming));
  CAB_CUDA(cudaEventRecord(net->aux_event[4], stream));
  for (int i = 0; i < 4; ++i) CAB_CUDA(cudaStreamWaitEvent(net->aux_stream[i], net->aux_event[4], 0));
  int k = 0;
  for (long b0 = 0; b0 < n; b0 += kPieceBoards, ++k) {
    int rc = piece(b0, std::min(kPieceBoards, n - b0), net->aux_stream[k & 3]);
    if (rc != CAB_OK) return rc;
  }
  for (int i = 0; i < 4; ++i) {
    CAB_CUDA(cudaEventRecord(net->aux_event[i], net->aux_stream[i]));
    CAB_CUDA(cudaStreamWaitEvent(stream, net->aux_event[i], 0));
  }
  return CAB_OK;
}

// bf16 tensor-core providers (cab_tc.cu): gradient and plain forward for nets whose hidden widths fit the GEMM tiles
int tc_grad(cab_net *net, const int8_t *d_codes, const double *d_targets, long n, double *d_grad, cudaStream_t stream);
int tc_forward_out(cab_net *net, const int8_t *d_codes, long n, double *d_out, cudaStream_t stream);
bool tc_fits(const cab_net *net);
int tc_turn_begin(cab_net *net, cudaStream_t stream);
int tc_turn_end(cab_net *net, cudaStream_t stream);

static int ensure_tout(cab_net *net, long n) {   // bf16 mode: only the output buffer of the re-forward
  if (n > net->tout_cap) {
    cudaFree(net->d_tout_tc);
    net->d_tout_tc = nullptr;
    CAB_CUDA(cudaMalloc(&net->d_tout_tc, n * sizeof(double)));
    net->tout_cap = n;
  }
  return CAB_OK;
}

// layout of a rank's exchange block (train_kernels.cuh: peer kernels)
static size_t peer_block_bytes(long nw, size_t *off_grad, size_t *off_e1, size_t *off_fg, size_t *off_fe) {
  const size_t g = (size_t) (nw + 4) * sizeof(double);
  off_grad[0] = 0; off_grad[1] = (g + 255) / 256 * 256;
  *off_e1 = off_grad[1] + (g + 255) / 256 * 256;
  *off_fg = *off_e1 + 2 * kPeerMax * sizeof(double);
  *off_fe = *off_fg + kPeerMax * sizeof(unsigned long long);
  return *off_fe + kPeerMax * sizeof(unsigned long long);
}
static PeerArgs peer_args(const cab_net *net, unsigned long long epoch) {
  PeerArgs a{};
  a.table = net->d_peer_table;
  a.rank = net->dp_rank; a.world = net->peer_world;
  a.epoch = epoch;
  a.nw = net->n_weights;
  peer_block_bytes(net->n_weights, a.off_grad, &a.off_e1, &a.off_flags_grad, &a.off_flags_e1);
  return a;
}

// One minibatch step, ENQUEUED on `stream` and left there: forward with saved activations, backward, gradient,
// (all-reduce), update with the device-resident lambda, re-forward, E', (all-reduce), decision + rollback + reset rule
// on the device (decide_kernel ...). Nothing here waits for the GPU: consecutive steps queue up behind each other.
static int enqueue_step(cab_net *net, const int8_t *d_codes, const double *d_targets, long n, cudaStream_t stream, bool bf16,
                        double *d_e_trace) {
  int rc = bf16 ? ensure_tout(net, n) : ensure_batch_buffers(net, n);
  if (rc != CAB_OK) return rc;
  double *d_tout = bf16 ? net->d_tout_tc : net->d_tout;
  TrainCtl *ctl = static_cast<TrainCtl *>(net->d_ctl);
  const long nw = net->n_weights;
  const int ew_grid = (int) std::min<long>((nw + 255) / 256, 148L * 8);
  const bool peer = net->peer_world > 1;
  PeerArgs pa{};
  if (peer) {   // this step's gradient buffer inside the exchange block (parity of the epoch)
    pa = peer_args(net, ++net->peer_epoch);
    net->d_grad = reinterpret_cast<double *>(static_cast<char *>(net->peer_block) + pa.off_grad[pa.epoch & 1]);
  }
  net->pack_dirty = net->tc_dirty = net->f16_dirty = true;     // the previous step may have rolled back or diluted the weights
  if (!bf16) {
    std::lock_guard<std::mutex> lk(net->mu);
    if ((rc = ensure_packed(net, stream, false)) != CAB_OK) return rc;
  }
  CAB_CUDA(cudaMemsetAsync(net->d_grad, 0, (nw + 4) * sizeof(double), stream));
  set_scalar_kernel<<<1, 1, 0, stream>>>(net->d_grad + nw + 2, (double) n);
  if (bf16) {
    if ((rc = tc_grad(net, d_codes, d_targets, n, net->d_grad, stream)) != CAB_OK) return rc;
  } else {
    // forward with saved activations (propagate_for_training), backward, gradient
    rc = fan_out(net, stream, n, [&](long b0, long nb, cudaStream_t s) {
      int r = launch_forward(net, d_codes + b0 * 64, nb, net->d_tout + b0, net->d_acts + b0 * 8, net->train_cap * 8, s);
      return r != CAB_OK ? r : launch_backward(net, d_targets, b0, nb, s);   // same stream: backward follows its forward
    });
    if (rc != CAB_OK) return rc;
    if ((rc = launch_grad(net, n, stream)) != CAB_OK) return rc;
  }
  const double *d_tail = net->d_grad + nw;      // (sum E, -, count) of the whole batch, for the decision
  if (peer) {   // all-reduce and update in one kernel over the ranks' mapped gradient buffers
    peer_signal_kernel<<<1, 32, 0, stream>>>(pa);
    peer_reduce_update_kernel<<<net->sm_count, 256, 0, stream>>>(pa, net->d_w, net->d_dw, ctl, net->d_scalars + 24);
    d_tail = net->d_scalars + 24;
    net->launches += 1;
  } else {
    if (net->nccl_comm != nullptr) {  // sum of (G, E, -, count) over ranks
      int nrc = g_nccl.AllReduce(net->d_grad, net->d_grad, (size_t) nw + 4, kNcclDouble, kNcclSum, net->nccl_comm, stream);
      if (nrc != 0) return nccl_fail(nrc, "ncclAllReduce(grad)");
    }
    apply_update_ctl_kernel<<<ew_grid, 256, 0, stream>>>(net->d_w, net->d_dw, net->d_grad, nw, ctl);
  }
  net->launches += 2;
  net->pack_dirty = net->tc_dirty = net->f16_dirty = true;
  // re-forward (network.cpp:265) and E'
  if (bf16) {
    if ((rc = tc_forward_out(net, d_codes, n, d_tout, stream)) != CAB_OK) return rc;
  } else {
    {
      std::lock_guard<std::mutex> lk(net->mu);
      if ((rc = ensure_packed(net, stream, false)) != CAB_OK) return rc;
    }
    rc = fan_out(net, stream, n, [&](long b0, long nb, cudaStream_t s) {
      return launch_forward(net, d_codes + b0 * 64, nb, net->d_tout + b0, nullptr, 0, s);
    });
    if (rc != CAB_OK) return rc;
  }
  CAB_CUDA(cudaMemsetAsync(net->d_scalars, 0, 2 * sizeof(double), stream));
  batch_error_kernel<<<(int) std::min<long>((n + 255) / 256, 148L * 4), 256, 0, stream>>>(d_tout, d_targets, n, net->d_scalars);
  if (peer) {
    peer_decide_kernel<<<1, 32, 0, stream>>>(pa, ctl, d_tail, net->d_scalars, d_e_trace);                   // E' over the ranks + the decision
  } else {
    if (net->nccl_comm != nullptr) {
      int nrc = g_nccl.AllReduce(net->d_scalars, net->d_scalars, 1, kNcclDouble, kNcclSum, net->nccl_comm, stream);
      if (nrc != 0) return nccl_fail(nrc, "ncclAllReduce(E')");
    }
    decide_kernel<<<1, 1, 0, stream>>>(ctl, d_tail, net->d_scalars, d_e_trace);                             // network.cpp:269-273, train.cpp:54-57
  }
  rollback_ctl_kernel<<<ew_grid, 256, 0, stream>>>(net->d_w, net->d_dw, nw, ctl);                           // :274-277 when rejected
  dropout_ctl_kernel<<<ew_grid, 256, 0, stream>>>(net->d_w, nw, 0.01, ctl);                                 // only behind a lambda reset
  net->launches += 4;
  net->pack_dirty = net->tc_dirty = net->f16_dirty = true;
  CAB_CUDA(cudaGetLastError());
  return CAB_OK;
}

static int upload_ctl(cab_net *net, double lambda, double max_const, double rate, int reset_rule, uint64_t seed, cudaStream_t stream) {
  TrainCtl h;
  memset(&h, 0, sizeof h);
  h.lambda = lambda; h.max_const = max_const; h.rate = rate; h.reset_rule = reset_rule; h.seed = seed;
  CAB_CUDA(cudaMemcpyAsync(net->d_ctl, &h, sizeof h, cudaMemcpyHostToDevice, stream));   // (pageable source: staged before the call returns)
  return CAB_OK;
}

// the one-step entry points: upload lambda, one step, wait, hand the decision back (Optimizer::optimize's signature)
static int train_batch_impl(cab_net *net, const int8_t *d_codes, const double *d_targets, long n, double *lambda,
                            double max_const, double rate, int *accepted, double *err, double *new_err,
                            cudaStream_t stream, bool bf16 = false) {
  if (!bf16 && !net->fp64_ok) { set_error("layer too wide for the fp64 training kernels"); return CAB_EINVAL; }
  if (bf16 && !tc_fits(net)) { set_error("bf16 tcgen05 training needs hidden widths that are multiples of 256"); return CAB_EINVAL; }
  int rc = ensure_train_state(net);
  if (rc != CAB_OK) return rc;
  std::unique_lock<std::mutex> turn(net->tc_mu, std::defer_lock);
  if (bf16) {   // the bf16 path's scratch is per net: this step takes its turn (cab_tc.cu)
    turn.lock();
    if ((rc = tc_turn_begin(net, stream)) != CAB_OK) return rc;
  }
  if ((rc = upload_ctl(net, *lambda, max_const, rate, 0, 0, stream)) != CAB_OK) return rc;
  if ((rc = enqueue_step(net, d_codes, d_targets, n, stream, bf16, nullptr)) != CAB_OK) return rc;
  TrainCtl h;
  CAB_CUDA(cudaMemcpyAsync(&h, net->d_ctl, sizeof h, cudaMemcpyDeviceToHost, stream));
  if (bf16 && (rc = tc_turn_end(net, stream)) != CAB_OK) return rc;
  CAB_CUDA(cudaStreamSynchronize(stream));
  if (h.peer_error) { set_error("data-parallel step: a peer's flag did not arrive (peer-memory mode)"); return CAB_ESTATE; }
  *lambda = h.lambda;
  if (accepted) *accepted = h.ok;
  if (err) *err = h.e0;
  if (new_err) *new_err = h.e1;
  return CAB_OK;
}

// n_total cases as consecutive minibatches of `batch` (the last one may be shorter), every step enqueued behind the
// previous one, ONE wait at the end: train_network's loop (train.cpp:39-68) over a resident shard without the host in it.
static int train_epoch_impl(cab_net *net, const int8_t *d_codes, const double *d_targets, long n_total, long batch, double *lambda,
                            double max_const, double rate, int reset_rule, uint64_t seed, cab_train_stats *stats, double *e_trace_host,
                            cudaStream_t stream, bool bf16) {
  if (!bf16 && !net->fp64_ok) { set_error("layer too wide for the fp64 training kernels"); return CAB_EINVAL; }
  if (bf16 && !tc_fits(net)) { set_error("bf16 tcgen05 training needs hidden widths that are multiples of 256"); return CAB_EINVAL; }
  int rc = ensure_train_state(net);
  if (rc != CAB_OK) return rc;
  std::unique_lock<std::mutex> turn(net->tc_mu, std::defer_lock);
  if (bf16) {
    turn.lock();
    if ((rc = tc_turn_begin(net, stream)) != CAB_OK) return rc;
  }
  const long steps = (n_total + batch - 1) / batch;
  struct DevBuf { double *p = nullptr; ~DevBuf() { cudaFree(p); } } trace;
  if (e_trace_host != nullptr) CAB_CUDA(cudaMalloc(&trace.p, (size_t) steps * 2 * sizeof(double)));
  if ((rc = upload_ctl(net, *lambda, max_const, rate, reset_rule, seed, stream)) != CAB_OK) return rc;
  for (long st = 0; st < steps; ++st) {
    const long b0 = st * batch, nb = std::min(batch, n_total - b0);
    if ((rc = enqueue_step(net, d_codes + b0 * 64, d_targets + b0, nb, stream, bf16, trace.p)) != CAB_OK) return rc;
  }
  TrainCtl h;
  CAB_CUDA(cudaMemcpyAsync(&h, net->d_ctl, sizeof h, cudaMemcpyDeviceToHost, stream));
  if (e_trace_host != nullptr) CAB_CUDA(cudaMemcpyAsync(e_trace_host, trace.p, (size_t) steps * 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
  if (bf16 && (rc = tc_turn_end(net, stream)) != CAB_OK) return rc;
  CAB_CUDA(cudaStreamSynchronize(stream));
  if (h.peer_error) { set_error("data-parallel epoch: a peer's flag did not arrive (peer-memory mode)"); return CAB_ESTATE; }
  *lambda = h.lambda;
  if (stats) {
    stats->steps = (long) h.steps; stats->accepted = (long) h.accepted; stats->resets = h.resets;
    stats->first_err = h.e_first; stats->last_err = h.e0; stats->last_new_err = h.e1; stats->last_accepted = h.ok;
  }
  return CAB_OK;
}

static int train_sequence_impl(cab_net *net, const int8_t *codes, const double *targets, long n, double *lambda,
                               double max_const, double rate, int reset_rule, uint64_t seed, uint8_t *accepted_host,
                               int *resets, double *err, double *new_err) {
  int rc = ensure_train_state(net);
  if (rc != CAB_OK) return rc;
  if ((rc = ensure_host_staging(net, n)) != CAB_OK) return rc;
  cudaStream_t s = net->train_stream;
  CAB_CUDA(cudaMemcpyAsync(net->d_tcodes, codes, n * 64, cudaMemcpyHostToDevice, s));
  CAB_CUDA(cudaMemcpyAsync(net->d_ttargets, targets, n * sizeof(double), cudaMemcpyHostToDevice, s));
  struct DevBytes {   // freed on every return path
    uint8_t *p = nullptr;
    ~DevBytes() { cudaFree(p); }
  } acc_buf;
  CAB_CUDA(cudaMalloc(&acc_buf.p, n));
  uint8_t *d_acc = acc_buf.p;
  SeqArgs a{};
  a.w = net->d_w; a.dw = net->d_dw;
  a.codes = net->d_tcodes; a.targets = net->d_ttargets;
  a.n_cases = n;
  a.n_layers = (int) net->dims.size();
  int off = 0;
  for (int l = 0; l < a.n_layers; ++l) {
    a.dims[l] = net->dims[l];
    a.noff[l] = off;
    off += net->dims[l];
    if (l + 1 < a.n_layers) a.woff[l] = net->woff[l];
  }
  a.noff[a.n_layers] = off;
  a.n_weights = net->n_weights;
  a.lambda_io = net->d_scalars + 8;
  a.max_const = max_const; a.rate = rate;
  a.apply_reset_rule = reset_rule;
  a.seed = seed;
  a.accepted = d_acc;
  a.resets = reinterpret_cast<int *>(net->d_scalars + 9);
  a.err_out = net->d_scalars + 10;
  CAB_CUDA(cudaMemcpyAsync(net->d_scalars + 8, lambda, sizeof(double), cudaMemcpyHostToDevice, s));
  const size_t smem = (size_t) 4 * off * sizeof(double);
  if (smem > 200 * 1024) { set_error("network too wide for the sequential trainer"); return CAB_EINVAL; }
  CAB_CUDA(cudaFuncSetAttribute(train_sequence_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  train_sequence_kernel<<<1, 1024, smem, s>>>(a);
  net->launches++;
  CAB_CUDA(cudaGetLastError());
  double h[4];
  CAB_CUDA(cudaMemcpyAsync(h, net->d_scalars + 8, 4 * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (accepted_host) CAB_CUDA(cudaMemcpyAsync(accepted_host, d_acc, n, cudaMemcpyDeviceToHost, s));
  CAB_CUDA(cudaStreamSynchronize(s));
  *lambda = h[0];
  if (resets) memcpy(resets, &h[1], sizeof(int));
  if (err) *err = h[2];
  if (new_err) *new_err = h[3];
  net->pack_dirty = net->tc_dirty = net->f16_dirty = true;
  return CAB_OK;
}

// leaves peer mode: closes the peers' mappings, frees the exchange block, d_grad back to the private buffer
void dp_peer_detach(cab_net *net) {
  if (net->peer_block == nullptr) return;
  cudaSetDevice(net->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < net->peer_world; ++r)
    if (r != net->dp_rank && net->peer_map[r] != nullptr) cudaIpcCloseMemHandle(net->peer_map[r]);
  cudaFree(net->peer_block);
  cudaFree(net->d_peer_table);
  net->peer_block = nullptr;
  net->d_peer_table = nullptr;
  for (auto &m : net->peer_map) m = nullptr;
  net->d_grad = net->d_grad_own;
  net->peer_world = 0;
  net->peer_epoch = 0;
}

}  // namespace cab

using namespace cab;

extern "C" {

int cab_dp_peer_export(cab_net *net, void *handle64_out) {
  if (!net || !handle64_out) { set_error("null argument"); return CAB_EINVAL; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI carries the IPC handle as 64 bytes");
  CAB_CUDA(cudaSetDevice(net->device));
  int rc = ensure_train_state(net);
  if (rc != CAB_OK) return rc;
  if (net->peer_block == nullptr) {
    size_t og[2], oe, ofg, ofe;
    const size_t bytes = peer_block_bytes(net->n_weights, og, &oe, &ofg, &ofe);
    CAB_CUDA(cudaMalloc(&net->peer_block, bytes));
    CAB_CUDA(memset_now(net->peer_block, 0, bytes));
    CAB_CUDA(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  CAB_CUDA(cudaIpcGetMemHandle(&h, net->peer_block));
  memcpy(handle64_out, &h, 64);
  return CAB_OK;
}

int cab_dp_peer_attach(cab_net *net, const void *handles64, int rank, int world) {
  if (!net || !handles64 || world < 2 || world > kPeerMax || rank < 0 || rank >= world) { set_error("bad argument (2 <= world <= %d)", kPeerMax); return CAB_EINVAL; }
  if (net->peer_block == nullptr) { set_error("cab_dp_peer_export first"); return CAB_ESTATE; }
  CAB_CUDA(cudaSetDevice(net->device));
  for (int r = 0; r < world; ++r) {
    if (r == rank) { net->peer_map[r] = net->peer_block; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char *>(handles64) + 64 * r, 64);
    void *p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      for (int q = 0; q < r; ++q)
        if (q != rank && net->peer_map[q]) { cudaIpcCloseMemHandle(net->peer_map[q]); net->peer_map[q] = nullptr; }
      return cuda_fail(e, "cudaIpcOpenMemHandle (peer access between the ranks' devices is required)", __FILE__, __LINE__);
    }
    net->peer_map[r] = p;
  }
  CAB_CUDA(cudaMalloc(&net->d_peer_table, kPeerMax * sizeof(void *)));
  CAB_CUDA(upload_now(net->d_peer_table, net->peer_map, kPeerMax * sizeof(void *)));
  net->d_grad_own = net->d_grad;
  net->dp_rank = rank;
  net->dp_world = world;
  net->peer_world = world;
  net->peer_epoch = 0;
  return CAB_OK;
}

int cab_dropout(cab_net *net, double rate, uint64_t seed, uint64_t offset, long *hits) {
  if (!net) { set_error("null argument"); return CAB_EINVAL; }
  CAB_CUDA(cudaSetDevice(net->device));
  CAB_CUDA(cudaDeviceSynchronize());
  unsigned long long *d_hits = nullptr;
  CAB_CUDA(cudaMalloc(&d_hits, sizeof(unsigned long long)));
  CAB_CUDA(memset_now(d_hits, 0, sizeof(unsigned long long)));
  dropout_kernel<<<(int) std::min<long>((net->n_weights + 255) / 256, 148L * 8), 256>>>(net->d_w, net->n_weights, rate,
                                                                                          seed, offset, d_hits);
  net->launches++;
  CAB_CUDA(cudaGetLastError());
  unsigned long long h = 0;
  CAB_CUDA(cudaMemcpy(&h, d_hits, sizeof h, cudaMemcpyDeviceToHost));
  cudaFree(d_hits);
  if (hits) *hits = (long) h;
  net->pack_dirty = net->tc_dirty = net->f16_dirty = true;
  return CAB_OK;
}

int cab_train_case(cab_net *net, const int8_t *codes64, double target, double *lambda, double max_const, double rate,
                   int *accepted, double *err, double *new_err) {
  if (!net || !codes64 || !lambda) { set_error("null argument"); return CAB_EINVAL; }
  CAB_CUDA(cudaSetDevice(net->device));
  uint8_t acc = 0;
  int rc = train_sequence_impl(net, codes64, &target, 1, lambda, max_const, rate, 0, 0, &acc, nullptr, err, new_err);
  if (rc == CAB_OK && accepted) *accepted = acc;
  return rc;
}

int cab_train_sequence(cab_net *net, const int8_t *codes, const double *targets, long n, double *lambda,
                       uint64_t dropout_seed, uint8_t *accepted_host, int *resets) {
  if (!net || !lambda || n < 0 || (n > 0 && (!codes || !targets))) { set_error("bad argument"); return CAB_EINVAL; }
  if (resets) *resets = 0;
  if (n == 0) return CAB_OK;
  CAB_CUDA(cudaSetDevice(net->device));
  return train_sequence_impl(net, codes, targets, n, lambda, 15.0, 1.1, 1, dropout_seed, accepted_host, resets, nullptr, nullptr);
}

int cab_train_batch_dev(cab_net *net, const int8_t *codes_dev, const double *targets_dev, long n, double *lambda,
                        double max_const, double rate, int *accepted, double *err, double *new_err, void *stream) {
  if (!net || !lambda || n <= 0 || !codes_dev || !targets_dev) { set_error("bad argument"); return CAB_EINVAL; }
  CAB_CUDA(cudaSetDevice(net->device));
  int rc = ensure_train_state(net);
  if (rc != CAB_OK) return rc;
  return train_batch_impl(net, codes_dev, targets_dev, n, lambda, max_const, rate, accepted, err, new_err,
                          stream ? (cudaStream_t) stream : net->train_stream);
}

int cab_train_batch_host(cab_net *net, const int8_t *codes, const double *targets, long n, double *lambda,
                         double max_const, double rate, int *accepted, double *err, double *new_err) {
  if (!net || !lambda || n <= 0 || !codes || !targets) { set_error("bad argument"); return CAB_EINVAL; }
  CAB_CUDA(cudaSetDevice(net->device));
  int rc = ensure_train_state(net);
  if (rc != CAB_OK) return rc;
  if ((rc = ensure_host_staging(net, n)) != CAB_OK) return rc;
  cudaStream_t s = net->train_stream;
  CAB_CUDA(cudaMemcpyAsync(net->d_tcodes, codes, n * 64, cudaMemcpyHostToDevice, s));
  CAB_CUDA(cudaMemcpyAsync(net->d_ttargets, targets, n * sizeof(double), cudaMemcpyHostToDevice, s));
  return train_batch_impl(net, net->d_tcodes, net->d_ttargets, n, lambda, max_const, rate, accepted, err, new_err, s);
}

int cab_train_epoch_dev(cab_net *net, const int8_t *codes_dev, const double *targets_dev, long n_cases, long batch, double *lambda,
                        double max_const, double rate, int reset_rule, uint64_t dropout_seed, cab_train_stats *stats, double *e_trace_host,
                        void *stream) {
  if (!net || !lambda || n_cases <= 0 || batch <= 0 || !codes_dev || !targets_dev) { set_error("bad argument"); return CAB_EINVAL; }
  CAB_CUDA(cudaSetDevice(net->device));
  int rc = ensure_train_state(net);
  if (rc != CAB_OK) return rc;
  return train_epoch_impl(net, codes_dev, targets_dev, n_cases, batch, lambda, max_const, rate, reset_rule, dropout_seed, stats, e_trace_host,
                          stream ? (cudaStream_t) stream : net->train_stream, false);
}

int cab_train_epoch_host(cab_net *net, const int8_t *codes_host, const double *targets_host, long n_cases, long batch, double *lambda,
                         double max_const, double rate, int reset_rule, uint64_t dropout_seed, cab_train_stats *stats, double *e_trace_host) {
  if (!net || !lambda || n_cases <= 0 || batch <= 0 || !codes_host || !targets_host) { set_error("bad argument"); return CAB_EINVAL; }
  CAB_CUDA(cudaSetDevice(net->device));
  int rc = ensure_train_state(net);
  if (rc != CAB_OK) return rc;
  if ((rc = ensure_host_staging(net, n_cases)) != CAB_OK) return rc;
  cudaStream_t s = net->train_stream;
  CAB_CUDA(cudaMemcpyAsync(net->d_tcodes, codes_host, (size_t) n_cases * 64, cudaMemcpyHostToDevice, s));
  CAB_CUDA(cudaMemcpyAsync(net->d_ttargets, targets_host, (size_t) n_cases * sizeof(double), cudaMemcpyHostToDevice, s));
  return train_epoch_impl(net, net->d_tcodes, net->d_ttargets, n_cases, batch, lambda, max_const, rate, reset_rule, dropout_seed, stats, e_trace_host,
                          s, false);
}

int cab_train_epoch_bf16_dev(cab_net *net, const int8_t *codes_dev, const double *targets_dev, long n_cases, long batch, double *lambda,
                             double max_const, double rate, int reset_rule, uint64_t dropout_seed, cab_train_stats *stats, double *e_trace_host,
                             void *stream) {
  if (!net || !lambda || n_cases <= 0 || batch <= 0 || !codes_dev || !targets_dev) { set_error("bad argument"); return CAB_EINVAL; }
  CAB_CUDA(cudaSetDevice(net->device));
  int rc = ensure_train_state(net);
  if (rc != CAB_OK) return rc;
  return train_epoch_impl(net, codes_dev, targets_dev, n_cases, batch, lambda, max_const, rate, reset_rule, dropout_seed, stats, e_trace_host,
                          stream ? (cudaStream_t) stream : net->train_stream, true);
}

int cab_train_batch_bf16_dev(cab_net *net, const int8_t *codes_dev, const double *targets_dev, long n, double *lambda,
                             double max_const, double rate, int *accepted, double *err, double *new_err, void *stream) {
  if (!net || !lambda || n <= 0 || !codes_dev || !targets_dev) { set_error("bad argument"); return CAB_EINVAL; }
  CAB_CUDA(cudaSetDevice(net->device));
  int rc = ensure_train_state(net);
  if (rc != CAB_OK) return rc;
  return train_batch_impl(net, codes_dev, targets_dev, n, lambda, max_const, rate, accepted, err, new_err,
                          stream ? (cudaStream_t) stream : net->train_stream, /*bf16=*/true);
}

int cab_train_batch_bf16_host(cab_net *net, const int8_t *codes, const double *targets, long n, double *lambda,
                              double max_const, double rate, int *accepted, double *err, double *new_err) {
  if (!net || !lambda || n <= 0 || !codes || !targets) { set_error("bad argument"); return CAB_EINVAL; }
  CAB_CUDA(cudaSetDevice(net->device));
  int rc = ensure_train_state(net);
  if (rc != CAB_OK) return rc;
  if ((rc = ensure_host_staging(net, n)) != CAB_OK) return rc;
  cudaStream_t s = net->train_stream;
  CAB_CUDA(cudaMemcpyAsync(net->d_tcodes, codes, n * 64, cudaMemcpyHostToDevice, s));
  CAB_CUDA(cudaMemcpyAsync(net->d_ttargets, targets, n * sizeof(double), cudaMemcpyHostToDevice, s));
  return train_batch_impl(net, net->d_tcodes, net->d_ttargets, n, lambda, max_const, rate, accepted, err, new_err, s, /*bf16=*/true);
}

// Forward (activations saved) + backward only: leaves sum_b dW (fp64, n,i,j order) in cab_grad_buffer()[0 .. nw) and
// sum_b E, -, count behind it. The timed unit of BASELINE.json configs[4] ("bf16 tcgen05 fwd+bwd").
int cab_grad_bf16_dev(cab_net *net, const int8_t *codes_dev, const double *targets_dev, long n, void *stream) {
  if (!net || n <= 0 || !codes_dev || !targets_dev) { set_error("bad argument"); return CAB_EINVAL; }
  CAB_CUDA(cudaSetDevice(net->device));
  if (!tc_fits(net)) { set_error("bf16 tcgen05 training needs hidden widths that are multiples of 256"); return CAB_EINVAL; }
  int rc = ensure_train_state(net);
  if (rc != CAB_OK) return rc;
  cudaStream_t s = stream ? (cudaStream_t) stream : net->train_stream;
  CAB_CUDA(cudaMemsetAsync(net->d_grad, 0, (net->n_weights + 4) * sizeof(double), s));
  set_scalar_kernel<<<1, 1, 0, s>>>(net->d_grad + net->n_weights + 2, (double) n);
  net->launches++;
  std::lock_guard<std::mutex> turn(net->tc_mu);
  if ((rc = tc_turn_begin(net, s)) != CAB_OK) return rc;
  if ((rc = tc_grad(net, codes_dev, targets_dev, n, net->d_grad, s)) != CAB_OK) return rc;
  return tc_turn_end(net, s);
}

int cab_dp_unique_id(void *id128_out) {
  if (!id128_out) { set_error("null argument"); return CAB_EINVAL; }
  int rc = load_nccl();
  if (rc != CAB_OK) return rc;
  int nrc = g_nccl.GetUniqueId(id128_out);
  return nrc == 0 ? CAB_OK : nccl_fail(nrc, "ncclGetUniqueId");
}

int cab_dp_init(cab_net *net, const void *id128, int rank, int world) {
  if (!net || !id128 || world < 1 || rank < 0 || rank >= world) { set_error("bad argument"); return CAB_EINVAL; }
  int rc = load_nccl();
  if (rc != CAB_OK) return rc;
  CAB_CUDA(cudaSetDevice(net->device));
  if ((rc = ensure_train_state(net)) != CAB_OK) return rc;
  Id128 id;
  memcpy(id.b, id128, 128);
  void *comm = nullptr;
  int nrc = g_nccl.CommInitRank(&comm, world, id, rank);
  if (nrc != 0) return nccl_fail(nrc, "ncclCommInitRank");
  net->nccl_comm = comm;
  net->dp_rank = rank;
  net->dp_world = world;
  return CAB_OK;
}

int cab_dp_shutdown(cab_net *net) {
  if (!net) return CAB_OK;
  drop_grad_plan(net);
  dp_peer_detach(net);
  if (net->nccl_comm != nullptr && g_nccl.CommDestroy) {
    g_nccl.CommDestroy(net->nccl_comm);
    net->nccl_comm = nullptr;
  }
  net->dp_world = 1;
  net->dp_rank = 0;
  return CAB_OK;
}

int cab_grad_buffer(cab_net *net, double **grad_dev, long *n_doubles) {
  if (!net || !grad_dev || !n_doubles) { set_error("null argument"); return CAB_EINVAL; }
  CAB_CUDA(cudaSetDevice(net->device));
  int rc = ensure_train_state(net);
  if (rc != CAB_OK) return rc;
  *grad_dev = net->d_grad;
  *n_doubles = net->n_weights + 4;
  return CAB_OK;
}

}  // extern "C"
