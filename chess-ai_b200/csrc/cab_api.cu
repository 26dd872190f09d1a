// cab_api.cu — implementation of include/chessai_b200.h: handle management, weight packing, encoder, forward
// entry points, network_dump text IO. Training lives in cab_train.cu. No CPU fallback anywhere: every compute
// entry point needs a CUDA device and says so when there is none.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "cab_internal.cuh"
#include "encode_simd.cuh"
#include "mlp_stream_kernel.cuh"

namespace cab {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  set_error("CUDA error '%s' in %s at %s:%d", cudaGetErrorString(e), what, file, line);
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return CAB_ENODEVICE;
  return CAB_ECUDA;
}

// ---- kernels ---------------------------------------------------------------------------------------------------

// K0 board encoder: piece records -> int8 codes. Integer-only, bit-exact with Piece::code()/0.4
// (/root/reference/src/chess/piece.cpp:103-105,145-147,179-181,260-262; type/colour tables piece.fwd.h:122-133,287-308).
// HBM-bound byte kernel: 192 B in, 64 B out per board. A CTA takes 64 boards (12 KB of records): the records arrive
// by three fully coalesced 16-byte loads per thread into shared memory, a thread then encodes 16 consecutive squares
// (48 B from shared memory at a 12-word stride: conflict-free for 128-bit accesses), four at a time in byte-parallel
// form (encode_simd.cuh: the scalar rule was issue-bound), and stores one 16-byte word.
constexpr int kEncThreads = 256;  // 16 squares per thread = 64 boards per CTA
__global__ void __launch_bounds__(kEncThreads) encode_kernel(const uint4 *__restrict__ rec, long n_hex,
                                                              uint4 *__restrict__ codes) {
  __shared__ uint4 sh[3 * kEncThreads];
  const long hex0 = (long) blockIdx.x * kEncThreads;            // first 16-square group of this CTA
  const long n_in = min((long) kEncThreads, n_hex - hex0) * 3;  // 16-byte words of records this CTA owns
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int idx = threadIdx.x + j * kEncThreads;
    if (idx < n_in) sh[idx] = __ldcs(rec + 3 * hex0 + idx);     // read once: streaming load
  }
  __syncthreads();
  if (hex0 + threadIdx.x >= n_hex) return;
  uint32_t w[12];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const uint4 v = sh[3 * threadIdx.x + j];
    w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
  }
  uint32_t o[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) o[q] = encode4(w[3 * q], w[3 * q + 1], w[3 * q + 2]);   // four squares = 12 bytes = words 3q .. 3q+2
  __stcs(codes + hex0 + threadIdx.x, make_uint4(o[0], o[1], o[2], o[3]));
}

// weights -> DMMA B-fragment stream (see cab_internal.cuh). One thread per packed double.
__global__ void __launch_bounds__(256) prepack_kernel(const double *__restrict__ w, double *__restrict__ pack,
                                                       const PackSeg *__restrict__ segs, int n_segs, long total) {
  for (long idx = blockIdx.x * (long) blockDim.x + threadIdx.x; idx < total; idx += (long) gridDim.x * blockDim.x) {
    int si = 0;
    while (si + 1 < n_segs && idx >= segs[si + 1].dst_off) ++si;
    const PackSeg sg = segs[si];
    const long local = idx - sg.dst_off;
    const int h = (int) (local & 1), lane = (int) ((local >> 1) & 31);
    const long rest = local >> 6;
    const int j = (int) (rest % sg.nb), t = (int) (rest / sg.nb);
    const int kin = 8 * t + 2 * (lane & 3) + h;          // contraction index (feature of the layer input side)
    const int nout = 8 * (sg.tile0 + j) + (lane >> 2);   // output feature
    double v = 0.0;
    if (!sg.transposed) {
      if (kin < sg.K && nout < sg.N) v = w[sg.w_off + (long) kin * sg.N + nout];
    } else {  // stream of W^T: contraction over N (kin indexes N), outputs over K (nout indexes K)
      if (kin < sg.N && nout < sg.K) v = w[sg.w_off + (long) nout * sg.N + kin];
    }
    pack[idx] = v;
  }
}

// Network::randomNetwork (network.cpp:167-173): U(-1,1) per weight; Philox counter = (w, stream), words 2,3 -> uniform
__global__ void __launch_bounds__(256) randomize_kernel(double *__restrict__ w, long n, uint64_t seed, uint64_t stream) {
  for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
    uint32_t r[4];
    philox4x32_10((uint32_t) i, (uint32_t) ((uint64_t) i >> 32), (uint32_t) stream, (uint32_t) (stream >> 32),
                  (uint32_t) seed, (uint32_t) (seed >> 32), r);
    w[i] = 2.0 * u53(r[2], r[3]) + -1.0;
  }
}

// ---- pipe-peak probes (roofline denominators that MEASURED_PEAKS.json lacks) -------------------------------------
__global__ void __launch_bounds__(256) peak_dmma_kernel(double *out, int iters, double seed) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) { c[i][0] = seed + i; c[i][1] = seed - i; }
  const double a = seed * 1e-3 + threadIdx.x * 1e-9, b = seed * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma884(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) peak_dfma_kernel(double *out, int iters, double seed) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = seed + i + threadIdx.x;
  const double a = seed * 1.0000001, b = seed * 0.5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) peak_ffma_kernel(float *out, int iters, float seed) {
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = seed + i + threadIdx.x;
  const float a = seed * 1.0000001f, b = seed * 0.5f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += acc[i];
  if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- host: stream plans ------------------------------------------------------------------------------------------

static int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Forward stream: layer l contracts over dims[l] (input units) and produces dims[l+1] (output tiles).
// Backward stream (transposed): weight layer l contracts over dims[l+1] and produces dims[l]; layers go top-down
// and layer 0 is skipped (omega of the input layer is never used, network.cpp:257 / SURVEY Q10).
static void build_plan(const cab_net *net, bool transposed, StreamPlan *plan) {
  plan->chunks.clear();
  plan->segs.clear();
  const int L = (int) net->dims.size();
  long off = 0;  // doubles
  bool single_block = true;
  int max_units = 8;
  std::vector<int> order;
  if (!transposed) for (int l = 0; l + 1 < L; ++l) order.push_back(l);
  else for (int l = L - 2; l >= 1; --l) order.push_back(l);
  for (int l : order) {
    const int in_w = transposed ? net->dims[l + 1] : net->dims[l], out_w = transposed ? net->dims[l] : net->dims[l + 1];
    if (ceil_div(out_w, 8) > kMaxNB) single_block = false;
    max_units = std::max(max_units, std::max(ceil_div(in_w, 8), ceil_div(out_w, 8)));
  }
  // activation buffer: in place when every layer fits one n-block, ping-pong otherwise
  plan->a_units = single_block ? max_units : 2 * max_units;
  int ping = 0;
  for (size_t oi = 0; oi < order.size(); ++oi) {
    const int l = order[oi];
    const int K = net->dims[l], N = net->dims[l + 1];
    const int in_w = transposed ? N : K, out_w = transposed ? K : N;
    const int t_in = ceil_div(in_w, 8), n_tiles = ceil_div(out_w, 8);
    const int n_blocks = ceil_div(n_tiles, kMaxNB);
    const int a_in = single_block ? 0 : ping * max_units, a_out_base = single_block ? 0 : (1 - ping) * max_units;
    int tile0 = 0;
    for (int b = 0; b < n_blocks; ++b) {
      const int nb = (n_tiles - tile0 + (n_blocks - b) - 1) / (n_blocks - b);  // balanced split
      PackSeg sg{};
      sg.dst_off = off;
      sg.w_off = net->woff[l];
      sg.K = K; sg.N = N;
      sg.t_in = t_in; sg.tile0 = tile0; sg.nb = nb; sg.transposed = transposed ? 1 : 0;
      sg.n_doubles = (long) t_in * nb * 64;
      plan->segs.push_back(sg);
      const int per_t_bytes = nb * kUnitBytes;
      const int tpc = std::max(1, net->stage_bytes / per_t_bytes);
      for (int t0 = 0; t0 < t_in; t0 += tpc) {
        ChunkDesc d{};
        d.tcount = (uint16_t) std::min(tpc, t_in - t0);
        d.t0 = (uint16_t) t0;
        d.src_off = (uint32_t) ((off + (long) t0 * nb * 64) * 8);
        d.bytes = (uint32_t) (d.tcount * per_t_bytes);
        d.tile0 = (uint16_t) tile0;
        d.nb = (uint8_t) nb;
        d.flags = 0;
        if (t0 == 0) d.flags |= kFirst;
        if (t0 + tpc >= t_in) d.flags |= kLast;
        if (oi + 1 == order.size() && b + 1 == n_blocks && (d.flags & kLast)) d.flags |= kFinal;
        d.a_in = (uint16_t) a_in;
        d.a_out = (uint16_t) (a_out_base + tile0);
        d.layer = (uint16_t) l;
        d.n_tiles = (uint16_t) n_tiles;
        d.save_off = transposed ? net->act_unit_off[l] : net->act_unit_off[l + 1];
        d.save_in = transposed ? net->act_unit_off[l + 1] : net->act_unit_off[l];
        plan->chunks.push_back(d);
      }
      off += sg.n_doubles;
      tile0 += nb;
    }
    ping = 1 - ping;
  }
  plan->total_doubles = off;
  // Forward stream: a final one-output layer (dims normalisation always appends 1) that fits one chunk, behind a layer
  // of a single n-block, is fused into that layer's epilogue as a dot product (mlp_stream_kernel.cuh: kDot).
  static const bool fuse = getenv("CAB_FWD_FUSE_OUT") ? atoi(getenv("CAB_FWD_FUSE_OUT")) != 0 : true;
  if (fuse && !transposed && L >= 3 && net->dims[L - 1] == 1 && plan->chunks.size() >= 2) {
    ChunkDesc &out_chunk = plan->chunks.back();
    ChunkDesc &prev_last = plan->chunks[plan->chunks.size() - 2];
    const int prev_tiles = ceil_div(net->dims[L - 2], 8);
    const bool one_chunk = (out_chunk.flags & kFirst) && (out_chunk.flags & kLast) && out_chunk.nb == 1 && out_chunk.tcount == prev_tiles;
    if (one_chunk && prev_tiles <= kMaxNB && prev_last.layer + 1 == out_chunk.layer && (prev_last.flags & kLast) && prev_last.nb == prev_tiles) {
      out_chunk.flags |= kDot;
      prev_last.flags |= kPreDot;
    }
  }
}

static int upload_plan(StreamPlan *plan) {
  if (plan->total_doubles == 0) return CAB_OK;
  CAB_CUDA(cudaMalloc(&plan->d_chunks, plan->chunks.size() * sizeof(ChunkDesc)));
  CAB_CUDA(upload_now(plan->d_chunks, plan->chunks.data(), plan->chunks.size() * sizeof(ChunkDesc)));
  {  // the kernel's start-up tables in one piece: the first kSmemChunks descriptors, then the exp table (the device's own copy)
    std::vector<uint8_t> blob(kBlobBytes, 0);
    memcpy(blob.data(), plan->chunks.data(), std::min(plan->chunks.size(), (size_t) kSmemChunks) * sizeof(ChunkDesc));
    CAB_CUDA(cudaMemcpyFromSymbol(blob.data() + kBlobChunkBytes, c_exp2_tbl, kExpTblSize * sizeof(double)));
    double *tbl8 = reinterpret_cast<double *>(blob.data() + kBlobChunkBytes);
    for (int j = 0; j < kExpTblSize; ++j) tbl8[j] *= 0.125;     // act_f_tbl wants T[j] / 8 (exact scaling)
    CAB_CUDA(cudaMalloc(&plan->d_blob, kBlobBytes));
    CAB_CUDA(upload_now(plan->d_blob, blob.data(), kBlobBytes));
  }
  CAB_CUDA(cudaMalloc(&plan->d_segs, plan->segs.size() * sizeof(PackSeg)));
  CAB_CUDA(upload_now(plan->d_segs, plan->segs.data(), plan->segs.size() * sizeof(PackSeg)));
  CAB_CUDA(cudaMalloc(&plan->d_pack, plan->total_doubles * sizeof(double)));
  return CAB_OK;
}

static void free_plan(StreamPlan *plan) {
  cudaFree(plan->d_chunks); cudaFree(plan->d_segs); cudaFree(plan->d_pack); cudaFree(plan->d_blob);
  plan->d_chunks = nullptr; plan->d_segs = nullptr; plan->d_pack = nullptr; plan->d_blob = nullptr;
}

size_t forward_smem_bytes(int a_units, int n_cons, int n_stages, int stage_bytes) {
  return fwd_smem_ring(n_stages, stage_bytes) + fwd_smem_abuf(n_cons, a_units) + fwd_smem_tail(n_stages);
}

void tc_free(cab_net *net);  // cab_tc.cu

constexpr size_t kSmemMax = 227 * 1024;
constexpr int kDefaultTeamShift = 3, kDefaultTeamLag = 0;   // teams of 8 warps; lag in cycles per team index

// CTA geometry: warp PAIRS per CTA (8 boards each), weight-ring depth, CTAs per SM. Wide layers trade pairs for
// activation space. n_cons counts pairs. CAB_FWD_PAIRS / CAB_FWD_STAGES override for experiments.
// n_boards > 0: the launch runs ALONE on the device (the caller serialises its launches), so it is spread over as many
// SMs as possible: pairs per CTA = ceil(8-board tiles / SMs). n_boards = 0: throughput geometry (other launches overlap).
Geometry choose_geometry(int a_units, int stage_bytes, long n_boards, int sm_count) {
  static const int env_w = getenv("CAB_FWD_PAIRS") ? atoi(getenv("CAB_FWD_PAIRS")) : 0;
  static const int env_s = getenv("CAB_FWD_STAGES") ? atoi(getenv("CAB_FWD_STAGES")) : 0;
  Geometry g{0, 0, 1};
  const int max_pairs = kConsumerWarps / 2;
  int want = env_w > 0 ? std::min(env_w, max_pairs) : max_pairs;
  if (env_w <= 0 && n_boards > 0) {
    const long tiles = (n_boards + kBoardsPerWarp - 1) / kBoardsPerWarp;
    want = (int) std::max<long>(1, std::min<long>(max_pairs, (tiles + sm_count - 1) / sm_count));
  }
  for (int p = want; p >= 1; --p) {
    if (forward_smem_bytes(a_units, p, 2, stage_bytes) > kSmemMax) continue;
    g.n_cons = p;
    static const int env_c = getenv("CAB_FWD_CTAS") ? atoi(getenv("CAB_FWD_CTAS")) : 0;
    const size_t budget = (env_c == 2 && p <= max_pairs / 2) ? kSmemMax / 2 - 1024 : kSmemMax;
    int st = env_s > 0 ? std::min(env_s, kMaxStages) : kMaxStages;
    while (st > 2 && forward_smem_bytes(a_units, p, st, stage_bytes) > budget) --st;
    g.n_stages = st;
    // two CTAs per SM only when both fit and together keep <= 16 warps
    g.ctas_per_sm = (p <= max_pairs / 2 && (forward_smem_bytes(a_units, p, st, stage_bytes) + 1024) * 2 <= kSmemMax) ? 2 : 1;
    return g;
  }
  return g;
}
int consumer_warps_for(int a_units, int stage_bytes) { return choose_geometry(a_units, stage_bytes, 0, 148).n_cons; }

// stagger of the warp teams inside a CTA (mlp_stream_kernel.cuh): cycles of delay per team index, teams = warp >> shift
static void team_stagger(int n_pairs, int *shift, int *lag) {
  static const int env_lag = getenv("CAB_FWD_LAG") ? atoi(getenv("CAB_FWD_LAG")) : -1;
  static const int env_shift = getenv("CAB_FWD_TEAM_SHIFT") ? atoi(getenv("CAB_FWD_TEAM_SHIFT")) : -1;
  *shift = env_shift >= 1 ? env_shift : kDefaultTeamShift;
  *lag = env_lag >= 0 ? env_lag : kDefaultTeamLag;
  if ((2 * n_pairs) >> *shift < 2) *lag = 0;   // a single team: nothing to stagger
}

// (re)build the packed streams from d_w on `stream`. Call with net->mu held. drain = true waits for the pack kernels
// before the flag falls, so evaluations on OTHER streams may follow; drain = false is for callers whose next launches
// are ordered after `stream` anyway (the trainer: same stream, or forked from it by event).
int ensure_packed(cab_net *net, cudaStream_t stream, bool drain) {
  if (!net->pack_dirty) return CAB_OK;
  for (StreamPlan *p : {&net->fwd, &net->bwd}) {
    if (p->total_doubles == 0) continue;
    const int grid = (int) std::min<long>((p->total_doubles + 255) / 256, 148L * 16);
    prepack_kernel<<<grid, 256, 0, stream>>>(net->d_w, p->d_pack, p->d_segs, (int) p->segs.size(), p->total_doubles);
    net->launches++;
  }
  CAB_CUDA(cudaGetLastError());
  if (drain) CAB_CUDA(cudaStreamSynchronize(stream));
  net->pack_dirty = false;
  return CAB_OK;
}

int launch_forward(cab_net *net, const int8_t *d_codes, long n_boards, double *d_out, double *d_save, long save_stride,
                   cudaStream_t stream, double *d_theta, bool alone) {
  if (n_boards <= 0) return CAB_OK;
  if (!net->fp64_ok) {
    set_error("layer too wide for the fp64 stream kernel (%d activation units per pair exceed shared memory); "
              "use the bf16 tcgen05 path (cab_eval_bf16_*)", net->fwd.a_units);
    return CAB_EINVAL;
  }
  ForwardArgs a{};
  a.pack = reinterpret_cast<const uint8_t *>(net->fwd.d_pack);
  a.chunks = net->fwd.d_chunks;
  a.n_chunks = (int) net->fwd.chunks.size();
  a.a_units = net->fwd.a_units;
  a.n_boards = n_boards;
  const Geometry geo = choose_geometry(a.a_units, net->stage_bytes, alone ? n_boards : 0, net->sm_count);
  a.stage_bytes = net->stage_bytes;
  const int n_cons = geo.n_cons;
  team_stagger(n_cons, &a.team_shift, &a.team_lag);
  a.n_stages = geo.n_stages;
  const int boards_per_cta = n_cons * kBoardsPerWarp;
  a.n_groups = (n_boards + boards_per_cta - 1) / boards_per_cta;
  a.codes = d_codes;
  a.out = d_out;
  a.save = d_save;
  a.save_theta = d_theta;
  a.save_stride = save_stride;
  a.blob = net->fwd.d_blob;
  for (int st = 0; st < geo.n_stages && st < kMaxStages; ++st) {   // the ring's first fills, as kernel parameters
    const ChunkDesc &cd = net->fwd.chunks[st % net->fwd.chunks.size()];
    a.fill_off[st] = cd.src_off;
    a.fill_bytes[st] = cd.bytes;
  }
  if (reinterpret_cast<uintptr_t>(d_codes) & 1) { set_error("board codes must be 2-byte aligned"); return CAB_EINVAL; }
  const size_t smem = forward_smem_bytes(a.a_units, n_cons, geo.n_stages, net->stage_bytes);
  // (one pass per CTA for a 4 096-board launch is deliberate: 2 / 4 passes per CTA with 6-16 streams measured 0.65 / 0.63 of
  // the pipe against 0.68 — CTAs that start at different moments desynchronise the SMs' requests for the weight stream)
  const int grid = (int) std::min<long>(a.n_groups, (long) net->sm_count * geo.ctas_per_sm);
  static const bool prof_env = getenv("CAB_FWD_PROFILE") != nullptr;
  if (prof_env) {
    // debug: per-warp phase cycle counters of ONE launch, printed to stderr (never on by default)
    long long *d_prof = nullptr;
    const long nw = (long) grid * n_cons * 2;
    CAB_CUDA(cudaMalloc(&d_prof, nw * 8 * sizeof(long long)));
    CAB_CUDA(memset_now(d_prof, 0, nw * 8 * sizeof(long long)));
    a.prof = d_prof;
    CAB_CUDA(cudaFuncSetAttribute(mlp_forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    mlp_forward_kernel<true><<<grid, n_cons * 64, smem, stream>>>(a);
    CAB_CUDA(cudaStreamSynchronize(stream));
    std::vector<long long> h(nw * 8);
    CAB_CUDA(cudaMemcpy(h.data(), d_prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(d_prof);
    double sum[8] = {0};
    for (long w = 0; w < nw; ++w) for (int i = 0; i < 8; ++i) sum[i] += (double) h[w * 8 + i];
    fprintf(stderr, "[cab profile] boards=%ld grid=%d pairs/cta=%d stages=%d  avg cycles per warp: load=%.0f desc=%.0f wait_full=%.0f "
            "mma=%.0f refill=%.0f act(incl pair barriers)=%.0f pair_barriers=%.0f total=%.0f\n", n_boards, grid, n_cons, geo.n_stages, sum[0] / nw, sum[1] / nw, sum[2] / nw,
            sum[3] / nw, sum[4] / nw, sum[5] / nw, sum[7] / nw, sum[6] / nw);
    net->launches++;
    return CAB_OK;
  }
  if (d_theta != nullptr) mlp_forward_kernel<false, true, true><<<grid, n_cons * 64, smem, stream>>>(a);
  else if (d_save != nullptr) mlp_forward_kernel<false, true><<<grid, n_cons * 64, smem, stream>>>(a);
  else mlp_forward_kernel<false><<<grid, n_cons * 64, smem, stream>>>(a);
  net->launches++;
  CAB_CUDA(cudaGetLastError());
  return CAB_OK;
}

static int select_device(int device) {
  static std::atomic<int> cached_count{-1};
  int n = cached_count.load();
  cudaError_t e = cudaSuccess;
  if (n <= 0) {
    e = cudaGetDeviceCount(&n);
    if (e == cudaSuccess && n > 0) cached_count.store(n);
  }
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device available (%s): libchessai_b200 has no CPU fallback", e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
    return CAB_ENODEVICE;
  }
  if (device < 0 || device >= n) {
    set_error("device %d out of range (have %d)", device, n);
    return CAB_EINVAL;
  }
  CAB_CUDA(cudaSetDevice(device));
  return CAB_OK;
}

static std::vector<int> normalize_dims(const int *dims, int n) {  // network.cpp:39-56
  std::vector<int> d;
  const std::vector<int> def = {64, 144, 225, 100, 1};
  if (dims == nullptr || n <= 0) return def;
  d.assign(dims, dims + n);
  if (d[0] != 64) d.insert(d.begin(), 64);
  if (d.back() != 1) d.push_back(1);
  if (d.size() <= 2) return def;
  return d;
}

static int init_net(cab_net *net, const std::vector<int> &dims, int device) {
  int rc = CAB_OK;
  cudaDeviceProp prop;
  CAB_CUDA(cudaGetDeviceProperties(&prop, device));
  net->sm_count = prop.multiProcessorCount;
  net->dims = dims;
  net->stage_bytes = getenv("CAB_STAGE_KB") ? std::max(16, std::min(64, atoi(getenv("CAB_STAGE_KB")))) * 1024 : kStageBytes;
  long w = 0;
  uint32_t u = 0;
  for (size_t l = 0; l < dims.size(); ++l) {
    net->act_unit_off.push_back(u);
    u += (uint32_t) ceil_div(dims[l], 8);
    if (l + 1 < dims.size()) { net->woff.push_back(w); w += (long) dims[l] * dims[l + 1]; }
  }
  net->act_units_total = u;
  net->n_weights = w;
  CAB_CUDA(cudaMalloc(&net->d_w, w * sizeof(double)));
  CAB_CUDA(memset_now(net->d_w, 0, w * sizeof(double)));
  build_plan(net, false, &net->fwd);
  build_plan(net, true, &net->bwd);
  if ((rc = upload_plan(&net->fwd)) != CAB_OK) return rc;
  if ((rc = upload_plan(&net->bwd)) != CAB_OK) return rc;
  // very wide layers do not fit the fp64 stream kernel's per-pair activation buffer: the network is still usable
  // through the bf16 tcgen05 path (cab_eval_bf16_*); the fp64 entry points then return CAB_EINVAL
  net->fp64_ok = consumer_warps_for(std::max(net->fwd.a_units, net->bwd.a_units), net->stage_bytes) > 0;
  CAB_CUDA(cudaFuncSetAttribute(mlp_forward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CAB_CUDA(cudaFuncSetAttribute(mlp_forward_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CAB_CUDA(cudaFuncSetAttribute(mlp_forward_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  net->pack_dirty = net->tc_dirty = net->f16_dirty = true;
  return CAB_OK;
}

static int create_exact(const std::vector<int> &dims, int device, cab_net **out) {
  for (int v : dims)
    if (v <= 0 || v > 8192) { set_error("layer width %d out of range [1, 8192]", v); return CAB_EINVAL; }
  if (dims.size() < 2 || dims.size() > 32) { set_error("bad layer count %zu", dims.size()); return CAB_EINVAL; }
  if (dims[0] != 64) { set_error("input layer must be 64 wide (one feature per square), got %d", dims[0]); return CAB_EINVAL; }
  int rc = select_device(device);
  if (rc != CAB_OK) return rc;
  cab_net *net = new cab_net();
  net->device = device;
  rc = init_net(net, dims, device);
  if (rc != CAB_OK) { cab_net_destroy(net); return rc; }   // a partially built net is released, not leaked
  *out = net;
  return CAB_OK;
}

EvalCtx *acquire_ctx(cab_net *net) {
  std::lock_guard<std::mutex> lk(net->mu);
  for (EvalCtx *c : net->ctxs)
    if (!c->busy) { c->busy = true; return c; }
  EvalCtx *c = new EvalCtx();
  c->busy = true;
  net->ctxs.push_back(c);
  return c;
}
void release_ctx(cab_net *net, EvalCtx *c) {
  std::lock_guard<std::mutex> lk(net->mu);
  c->busy = false;
}

constexpr long kEvalChunkBoards = 8192;   // measured best (4096: 1.72e8, 8192: 1.80e8, 16384: 1.75e8, 32768: 1.73e8)  // host<->device pipelining granularity of the *_host entry points

static int ctx_reserve(EvalCtx *c, long boards, bool records) {
  for (int i = 0; i < kCtxStreams; ++i)
    if (c->stream[i] == nullptr) CAB_CUDA(cudaStreamCreateWithFlags(&c->stream[i], cudaStreamNonBlocking));
  if (boards > c->cap_boards) {
    for (int i = 0; i < kCtxStreams; ++i) {
      cudaFree(c->d_codes[i]); cudaFree(c->d_out[i]); cudaFree(c->d_records[i]);
      c->d_records[i] = nullptr;
      CAB_CUDA(cudaMalloc(&c->d_codes[i], boards * 64));
      CAB_CUDA(cudaMalloc(&c->d_out[i], boards * sizeof(double)));
    }
    c->cap_boards = boards;
  }
  if (records)
    for (int i = 0; i < kCtxStreams; ++i)
      if (c->d_records[i] == nullptr) CAB_CUDA(cudaMalloc(&c->d_records[i], c->cap_boards * 192));
  return CAB_OK;
}

static int eval_host_impl(cab_net *net, const int8_t *codes, const uint8_t *records, long n, double *out) {
  int rc = select_device(net->device);
  if (rc != CAB_OK) return rc;
  EvalCtx *c = acquire_ctx(net);
  static const long env_chunk = getenv("CAB_EVAL_CHUNK") ? atol(getenv("CAB_EVAL_CHUNK")) : 0;
  const long chunk = std::min(n, env_chunk > 0 ? env_chunk : kEvalChunkBoards);
  rc = ctx_reserve(c, chunk, records != nullptr);
  if (rc == CAB_OK) {
    {
      std::lock_guard<std::mutex> lk(net->mu);
      rc = ensure_packed(net, c->stream[0], true);
    }
    int i = 0;
    for (long b0 = 0; b0 < n && rc == CAB_OK; b0 += chunk, i = (i + 1) % kCtxStreams) {
      const long nb = std::min(chunk, n - b0);
      cudaStream_t s = c->stream[i];
      if (records != nullptr) {
        if (cudaMemcpyAsync(c->d_records[i], records + b0 * 192, nb * 192, cudaMemcpyHostToDevice, s) != cudaSuccess) { rc = CAB_ECUDA; break; }
        rc = cab_encode_dev(c->d_records[i], nb, c->d_codes[i], s);
        net->launches++;
      } else {
        if (cudaMemcpyAsync(c->d_codes[i], codes + b0 * 64, nb * 64, cudaMemcpyHostToDevice, s) != cudaSuccess) { rc = CAB_ECUDA; break; }
      }
      if (rc == CAB_OK) rc = launch_forward(net, c->d_codes[i], nb, c->d_out[i], nullptr, 0, s);
      if (rc == CAB_OK && cudaMemcpyAsync(out + b0, c->d_out[i], nb * sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = CAB_ECUDA;
    }
    for (int k = 0; k < kCtxStreams; ++k)
      if (cudaStreamSynchronize(c->stream[k]) != cudaSuccess && rc == CAB_OK) rc = CAB_ECUDA;
    if (rc == CAB_ECUDA) set_error("CUDA failure in eval: %s", cudaGetErrorString(cudaGetLastError()));
  }
  release_ctx(net, c);
  return rc;
}

}  // namespace cab

using namespace cab;

// ================================================================================================================
extern "C" {

int cab_version(void) { return 1; }
const char *cab_last_error(void) { return g_err; }

int cab_device_count(int *count) {
  if (!count) { set_error("null argument"); return CAB_EINVAL; }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { *count = 0; return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__); }
  *count = n;
  return CAB_OK;
}

int cab_net_create(const int *dims, int n_dims, int device, cab_net **out) {
  if (!out) { set_error("null argument"); return CAB_EINVAL; }
  *out = nullptr;
  return create_exact(normalize_dims(dims, n_dims), device, out);
}

int cab_net_destroy(cab_net *net) {
  if (!net) return CAB_OK;
  cudaSetDevice(net->device);
  cab_dp_shutdown(net);
  tc_free(net);
  for (EvalCtx *c : net->ctxs) {
    for (int i = 0; i < kCtxStreams; ++i) {
      if (c->stream[i]) cudaStreamDestroy(c->stream[i]);
      cudaFree(c->d_codes[i]); cudaFree(c->d_out[i]); cudaFree(c->d_records[i]);
    }
    delete c;
  }
  free_plan(&net->fwd); free_plan(&net->bwd);
  dp_peer_detach(net);   // d_grad points into the peer exchange block while attached
  cudaFree(net->d_w); cudaFree(net->d_dw); cudaFree(net->d_grad); cudaFree(net->d_acts); cudaFree(net->d_psis);
  cudaFree(net->d_tcodes); cudaFree(net->d_ttargets); cudaFree(net->d_tout); cudaFree(net->d_tout_tc); cudaFree(net->d_scalars); cudaFree(net->d_ctl);
  if (net->train_stream) cudaStreamDestroy(net->train_stream);
  for (int i = 0; i < 4; ++i) if (net->aux_stream[i]) cudaStreamDestroy(net->aux_stream[i]);
  for (int i = 0; i < 5; ++i) if (net->aux_event[i]) cudaEventDestroy(net->aux_event[i]);
  delete net;
  return CAB_OK;
}

int cab_net_clone(const cab_net *net, cab_net **out) {  // Network::clone — weights only (network.cpp:195-205)
  if (!net || !out) { set_error("null argument"); return CAB_EINVAL; }
  int rc = create_exact(net->dims, net->device, out);
  if (rc != CAB_OK) return rc;
  CAB_CUDA(cudaMemcpy((*out)->d_w, net->d_w, net->n_weights * sizeof(double), cudaMemcpyDeviceToDevice));
  (*out)->pack_dirty = true;
  return CAB_OK;
}

int cab_net_num_layers(const cab_net *net, int *n) {
  if (!net || !n) { set_error("null argument"); return CAB_EINVAL; }
  *n = (int) net->dims.size();
  return CAB_OK;
}
int cab_net_dims(const cab_net *net, int *dims_out) {
  if (!net || !dims_out) { set_error("null argument"); return CAB_EINVAL; }
  for (size_t i = 0; i < net->dims.size(); ++i) dims_out[i] = net->dims[i];
  return CAB_OK;
}
int cab_net_num_weights(const cab_net *net, long *n) {
  if (!net || !n) { set_error("null argument"); return CAB_EINVAL; }
  *n = net->n_weights;
  return CAB_OK;
}
int cab_net_device(const cab_net *net, int *device) {
  if (!net || !device) { set_error("null argument"); return CAB_EINVAL; }
  *device = net->device;
  return CAB_OK;
}

int cab_net_set_weights(cab_net *net, const double *w, long n) {
  if (!net || !w) { set_error("null argument"); return CAB_EINVAL; }
  if (n != net->n_weights) { set_error("expected %ld weights, got %ld", net->n_weights, n); return CAB_EINVAL; }
  int rc = select_device(net->device);
  if (rc != CAB_OK) return rc;
  CAB_CUDA(cudaDeviceSynchronize());
  CAB_CUDA(upload_now(net->d_w, w, n * sizeof(double)));
  net->pack_dirty = net->tc_dirty = net->f16_dirty = true;
  return CAB_OK;
}

int cab_net_get_weights(const cab_net *net, double *w, long n) {
  if (!net || !w) { set_error("null argument"); return CAB_EINVAL; }
  if (n != net->n_weights) { set_error("expected %ld weights, got %ld", net->n_weights, n); return CAB_EINVAL; }
  int rc = select_device(net->device);
  if (rc != CAB_OK) return rc;
  CAB_CUDA(cudaDeviceSynchronize());
  CAB_CUDA(cudaMemcpy(w, net->d_w, n * sizeof(double), cudaMemcpyDeviceToHost));
  return CAB_OK;
}

int cab_net_randomize(cab_net *net, uint64_t seed, uint64_t stream) {
  if (!net) { set_error("null argument"); return CAB_EINVAL; }
  int rc = select_device(net->device);
  if (rc != CAB_OK) return rc;
  randomize_kernel<<<(int) std::min<long>((net->n_weights + 255) / 256, 148L * 8), 256>>>(net->d_w, net->n_weights, seed, stream);
  net->launches++;
  CAB_CUDA(cudaGetLastError());
  CAB_CUDA(cudaDeviceSynchronize());
  net->pack_dirty = net->tc_dirty = net->f16_dirty = true;
  return CAB_OK;
}

int cab_net_zero(cab_net *net) {
  if (!net) { set_error("null argument"); return CAB_EINVAL; }
  int rc = select_device(net->device);
  if (rc != CAB_OK) return rc;
  CAB_CUDA(cudaDeviceSynchronize());
  CAB_CUDA(memset_now(net->d_w, 0, net->n_weights * sizeof(double)));
  net->pack_dirty = net->tc_dirty = net->f16_dirty = true;
  return CAB_OK;
}

// ---- network_dump text ------------------------------------------------------------------------------------------
int cab_net_load_text(const char *path, int device, cab_net **out) {
  if (!path || !out) { set_error("null argument"); return CAB_EINVAL; }
  *out = nullptr;
  FILE *f = fopen(path, "rb");
  if (!f) { set_error("cannot open network dump '%s'", path); return CAB_EIO; }
  int L = 0;
  if (fscanf(f, "%d", &L) != 1 || L < 2 || L > 32) { fclose(f); set_error("'%s': bad layer count", path); return CAB_EIO; }
  std::vector<int> dims(L);
  for (int l = 0; l < L; ++l)
    if (fscanf(f, "%d", &dims[l]) != 1) { fclose(f); set_error("'%s': bad dimension list", path); return CAB_EIO; }
  long nw = 0;
  for (int l = 0; l + 1 < L; ++l) nw += (long) dims[l] * dims[l + 1];
  if (nw <= 0 || nw > (1L << 31)) { fclose(f); set_error("'%s': implausible weight count %ld", path, nw); return CAB_EIO; }
  std::vector<double> w(nw);
  // whitespace-delimited tokens, like operator>> (network.cpp:401-404); strtod on a slurped buffer is ~10x faster than fscanf
  long pos = ftell(f);
  fseek(f, 0, SEEK_END);
  long end = ftell(f);
  fseek(f, pos, SEEK_SET);
  std::vector<char> buf(end - pos + 1);
  size_t got = fread(buf.data(), 1, end - pos, f);
  fclose(f);
  buf[got] = 0;
  char *p = buf.data();
  for (long i = 0; i < nw; ++i) {
    char *q = nullptr;
    w[i] = strtod(p, &q);
    if (q == p) { set_error("'%s': expected %ld weights, found %ld", path, nw, i); return CAB_EIO; }
    p = q;
  }
  int rc = create_exact(dims, device, out);
  if (rc != CAB_OK) return rc;
  rc = cab_net_set_weights(*out, w.data(), nw);
  if (rc != CAB_OK) { cab_net_destroy(*out); *out = nullptr; }
  return rc;
}

int cab_net_dump_text(const cab_net *net, const char *path) {
  if (!net || !path) { set_error("null argument"); return CAB_EINVAL; }
  std::vector<double> w(net->n_weights);
  int rc = cab_net_get_weights(net, w.data(), net->n_weights);
  if (rc != CAB_OK) return rc;
  FILE *f = fopen(path, "wb");
  if (!f) { set_error("cannot open '%s' for writing", path); return CAB_EIO; }
  const int L = (int) net->dims.size();
  fprintf(f, "%d\n", L);                                      // network.cpp:411
  for (int l = 0; l + 1 < L; ++l) fprintf(f, "%d ", net->dims[l]);
  fprintf(f, "%d\n", net->dims[L - 1]);                       // network.cpp:414-416
  for (long i = 0; i < net->n_weights; ++i) fprintf(f, "%.17e ", w[i]);  // string::from_double + " " (network.cpp:422)
  if (fclose(f) != 0) { set_error("write to '%s' failed", path); return CAB_EIO; }
  return CAB_OK;
}

// ---- encoder ----------------------------------------------------------------------------------------------------
int cab_encode_dev(const uint8_t *records_dev, long n_boards, int8_t *codes_dev, void *stream) {
  if (n_boards < 0 || (n_boards > 0 && (!records_dev || !codes_dev))) { set_error("bad argument"); return CAB_EINVAL; }
  if (n_boards == 0) return CAB_OK;
  const long n_hex = n_boards * 4;  // groups of 16 squares
  if ((reinterpret_cast<uintptr_t>(records_dev) | reinterpret_cast<uintptr_t>(codes_dev)) & 15) {
    set_error("encoder buffers must be 16-byte aligned");
    return CAB_EINVAL;
  }
  const long grid = (n_hex + kEncThreads - 1) / kEncThreads;
  if (grid > 0x7fffffffL) { set_error("too many boards for one encoder launch"); return CAB_EINVAL; }
  encode_kernel<<<(unsigned) grid, kEncThreads, 0, (cudaStream_t) stream>>>(
      reinterpret_cast<const uint4 *>(records_dev), n_hex, reinterpret_cast<uint4 *>(codes_dev));
  CAB_CUDA(cudaGetLastError());
  return CAB_OK;
}

int cab_encode_host(int device, const uint8_t *records, long n, int8_t *codes) {
  if (n < 0 || (n > 0 && (!records || !codes))) { set_error("bad argument"); return CAB_EINVAL; }
  int rc = select_device(device);
  if (rc != CAB_OK || n == 0) return rc;
  uint8_t *d_r = nullptr;
  int8_t *d_c = nullptr;
  CAB_CUDA(cudaMalloc(&d_r, n * 192));
  CAB_CUDA(cudaMalloc(&d_c, n * 64));
  CAB_CUDA(upload_now(d_r, records, n * 192));
  rc = cab_encode_dev(d_r, n, d_c, nullptr);
  if (rc == CAB_OK && cudaMemcpy(codes, d_c, n * 64, cudaMemcpyDeviceToHost) != cudaSuccess) rc = CAB_ECUDA;
  cudaFree(d_r); cudaFree(d_c);
  return rc;
}

// ---- forward ----------------------------------------------------------------------------------------------------
int cab_eval_dev(cab_net *net, const int8_t *codes_dev, long n, double *out_dev, void *stream) {
  if (!net || n < 0 || (n > 0 && (!codes_dev || !out_dev))) { set_error("bad argument"); return CAB_EINVAL; }
  int rc = select_device(net->device);
  if (rc != CAB_OK) return rc;
  {
    std::lock_guard<std::mutex> lk(net->mu);
    if ((rc = ensure_packed(net, (cudaStream_t) stream, true)) != CAB_OK) return rc;  // other streams may evaluate next
  }
  // one launch on the caller's stream: spread it over the whole device (latency geometry)
  return launch_forward(net, codes_dev, n, out_dev, nullptr, 0, (cudaStream_t) stream, nullptr, /*alone=*/true);
}

int cab_eval_dev_batches(cab_net *net, const int8_t *codes_dev, long n, long batch, double *out_dev,
                         void *const *streams, int n_streams) {
  if (!net || n < 0 || batch <= 0 || n_streams <= 0 || !streams || (n > 0 && (!codes_dev || !out_dev))) {
    set_error("bad argument");
    return CAB_EINVAL;
  }
  int rc = select_device(net->device);
  if (rc != CAB_OK) return rc;
  {
    std::lock_guard<std::mutex> lk(net->mu);
    if ((rc = ensure_packed(net, (cudaStream_t) streams[0], true)) != CAB_OK) return rc;
  }
  int i = 0;
  for (long b0 = 0; b0 < n; b0 += batch, ++i) {
    // a single stream serialises the launches, each then has the device to itself: latency geometry
    rc = launch_forward(net, codes_dev + b0 * 64, std::min(batch, n - b0), out_dev + b0, nullptr, 0,
                        (cudaStream_t) streams[i % n_streams], nullptr, /*alone=*/n_streams == 1);
    if (rc != CAB_OK) return rc;
  }
  return CAB_OK;
}

int cab_eval_host(cab_net *net, const int8_t *codes, long n, double *out) {
  if (!net || n < 0 || (n > 0 && (!codes || !out))) { set_error("bad argument"); return CAB_EINVAL; }
  if (n == 0) return CAB_OK;
  return eval_host_impl(net, codes, nullptr, n, out);
}

int cab_eval_records_host(cab_net *net, const uint8_t *records, long n, double *out) {
  if (!net || n < 0 || (n > 0 && (!records || !out))) { set_error("bad argument"); return CAB_EINVAL; }
  if (n == 0) return CAB_OK;
  return eval_host_impl(net, nullptr, records, n, out);
}

static int forward_full_impl(cab_net *net, const int8_t *codes, long n, double *acts_host, double *thetas_host) {
  int rc = select_device(net->device);
  if (rc != CAB_OK) return rc;
  const long pad = (n + 63) / 64 * 64;
  const long stride = pad * 8;
  const size_t tensor = (size_t) net->act_units_total * stride;
  struct Dev {   // released on every return path
    void *p = nullptr;
    ~Dev() { cudaFree(p); }
  } codes_b, out_b, save_b, theta_b;
  CAB_CUDA(cudaMalloc(&codes_b.p, n * 64));
  CAB_CUDA(cudaMalloc(&out_b.p, n * sizeof(double)));
  CAB_CUDA(cudaMalloc(&save_b.p, tensor * sizeof(double)));
  CAB_CUDA(memset_now(save_b.p, 0, tensor * sizeof(double)));
  if (thetas_host != nullptr) {
    CAB_CUDA(cudaMalloc(&theta_b.p, tensor * sizeof(double)));
    CAB_CUDA(memset_now(theta_b.p, 0, tensor * sizeof(double)));
  }
  CAB_CUDA(upload_now(codes_b.p, codes, n * 64));
  {
    std::lock_guard<std::mutex> lk(net->mu);
    rc = ensure_packed(net, nullptr, true);
  }
  if (rc != CAB_OK) return rc;
  rc = launch_forward(net, (const int8_t *) codes_b.p, n, (double *) out_b.p, (double *) save_b.p, stride, nullptr, (double *) theta_b.p);
  if (rc != CAB_OK) return rc;
  long row = 0;
  for (size_t l = 1; l < net->dims.size(); ++l) row += net->dims[l];
  std::vector<double> h(tensor);
  // unit layout [unit][board][8] -> board-major rows of sum(dims[1..]) doubles
  auto gather = [&](const void *dev, double *host) -> int {
    CAB_CUDA(cudaMemcpy(h.data(), dev, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (long b = 0; b < n; ++b) {
      long o = 0;
      for (size_t l = 1; l < net->dims.size(); ++l)
        for (int j = 0; j < net->dims[l]; ++j, ++o)
          host[b * row + o] = h[(size_t) (net->act_unit_off[l] + j / 8) * stride + b * 8 + (j % 8)];
    }
    return CAB_OK;
  };
  if (acts_host != nullptr && (rc = gather(save_b.p, acts_host)) != CAB_OK) return rc;
  if (thetas_host != nullptr && (rc = gather(theta_b.p, thetas_host)) != CAB_OK) return rc;
  return CAB_OK;
}

int cab_forward_full_host(cab_net *net, const int8_t *codes, long n, double *acts_host) {
  if (!net || n < 0 || (n > 0 && (!codes || !acts_host))) { set_error("bad argument"); return CAB_EINVAL; }
  if (n == 0) return CAB_OK;
  return forward_full_impl(net, codes, n, acts_host, nullptr);
}

int cab_forward_thetas_host(cab_net *net, const int8_t *codes, long n, double *acts_host, double *thetas_host) {
  if (!net || n < 0 || (n > 0 && (!codes || !thetas_host))) { set_error("bad argument"); return CAB_EINVAL; }
  if (n == 0) return CAB_OK;
  return forward_full_impl(net, codes, n, acts_host, thetas_host);
}

// ---- measurement helpers ----------------------------------------------------------------------------------------
int cab_measure_pipe_peak(int device, int kind, double *tflops) {
  if (!tflops || kind < 0 || kind > 2) { set_error("bad argument"); return CAB_EINVAL; }
  int rc = select_device(device);
  if (rc != CAB_OK) return rc;
  cudaDeviceProp prop;
  CAB_CUDA(cudaGetDeviceProperties(&prop, device));
  const int grid = prop.multiProcessorCount * 2, iters = 4096;
  double *d = nullptr;
  CAB_CUDA(cudaMalloc(&d, (size_t) grid * 256 * sizeof(double)));
  cudaEvent_t e0, e1;
  CAB_CUDA(cudaEventCreate(&e0));
  CAB_CUDA(cudaEventCreate(&e1));
  auto launch = [&] {
    if (kind == 0) peak_dmma_kernel<<<grid, 256>>>(d, iters, 1.0);
    else if (kind == 1) peak_dfma_kernel<<<grid, 256>>>(d, iters, 1.0);
    else peak_ffma_kernel<<<grid, 256>>>((float *) d, iters, 1.0f);
  };
  for (int i = 0; i < 3; ++i) launch();
  CAB_CUDA(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 10; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    CAB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  double flop;
  if (kind == 0) flop = 2.0 * 256 * 16 * iters * (double) grid * 8;
  else if (kind == 1) flop = 2.0 * 16 * iters * (double) grid * 256;
  else flop = 2.0 * 32 * iters * (double) grid * 256;
  *tflops = flop / best * 1e-9;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  CAB_CUDA(cudaGetLastError());
  return CAB_OK;
}

int cab_net_launch_count(const cab_net *net, long *launches) {
  if (!net || !launches) { set_error("null argument"); return CAB_EINVAL; }
  *launches = net->launches.load();
  return CAB_OK;
}

}  // extern "C"
