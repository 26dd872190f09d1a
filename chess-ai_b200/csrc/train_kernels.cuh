// train_kernels.cuh — CUDA kernels of the training path (K2-K5).
//
// Reference: network::Optimizer::optimize / dropout (/root/reference/src/mcts_network/network.cpp:223-292) and the
// train loop of /root/reference/src/main/network/train.cpp:39-68.
//
//  train_sequence_kernel   the reference step exactly as written, one case after another, in ONE persistent CTA:
//                          same summation order (ascending i per output, ascending j per omega), separate multiply
//                          and add (no FMA contraction), (psi*lambda)*a update, accept / reject with rollback, the
//                          lambda reset + Philox dropout rule. This is the B == 1 parity path.
//  mlp_backward_kernel     minibatch: psi/omega recurrences for 8 boards per warp pair on the fp64 tensor pipe, fed
//                          by the packed W^T stream (same ring / pairing as the forward kernel); saves psi.
//  grad_kernel             minibatch: G_l += A_l^T Psi_{l+1} (contraction over boards) on the fp64 tensor pipe.
//  apply / rollback / error kernels: elementwise.
#pragma once
#include "cab_internal.cuh"
#include "mlp_stream_kernel.cuh"

namespace cab {

// ---------------------------------------------------------------------------------------------------------------
// B == 1, exact order
// ---------------------------------------------------------------------------------------------------------------
struct SeqArgs {
  double *w, *dw;            // flat weights / last update, n,i,j order
  const int8_t *codes;       // [n_cases][64]
  const double *targets;     // [n_cases]
  long n_cases;
  int n_layers;
  int dims[32];
  long woff[32];
  int noff[33];
  long n_weights;
  double *lambda_io;         // in: starting lambda, out: final lambda
  double max_const, rate;
  int apply_reset_rule;      // train.cpp:54-57 (lambda reset + dropout) after every step
  unsigned long long seed;   // Philox key of the dropout
  uint8_t *accepted;         // [n_cases] or nullptr
  int *resets;               // out
  double *err_out;           // [2]: E, E' of the LAST step
};

// forward over all layers on shared-memory neuron arrays; thread j owns output j (ascending-i sum, mul then add)
__device__ __forceinline__ void seq_forward(const SeqArgs &a, double *neurons, double *thetas, bool store_thetas) {
  for (int l = 1; l < a.n_layers; ++l) {
    const int K = a.dims[l - 1], N = a.dims[l];
    const double *W = a.w + a.woff[l - 1];
    const double *in = neurons + a.noff[l - 1];
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      double sum = 0.0;
#pragma unroll 4
      for (int i = 0; i < K; ++i) sum = __dadd_rn(sum, __dmul_rn(in[i], W[(long) i * N + j]));  // network.cpp:108-109
      if (store_thetas) thetas[a.noff[l] + j] = sum;                                             // :150
      neurons[a.noff[l] + j] = act_f(sum);                                                       // :112 / :151
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024, 1) train_sequence_kernel(const SeqArgs a) {
  extern __shared__ double sm[];
  const int nn = a.noff[a.n_layers];
  double *neurons = sm, *thetas = sm + nn, *omegas = sm + 2 * nn, *psis = sm + 3 * nn;
  __shared__ double s_lambda, s_err, s_new_err;
  __shared__ int s_accept, s_resets;
  const int out = a.n_layers - 1;
  for (int i = threadIdx.x; i < nn; i += blockDim.x) omegas[i] = 0.0;
  if (threadIdx.x == 0) { s_lambda = *a.lambda_io; s_resets = 0; s_err = 0; s_new_err = 0; }
  __syncthreads();

  for (long c = 0; c < a.n_cases; ++c) {
    // loadBoard (network.cpp:119-123 via :228)
    for (int i = threadIdx.x; i < 64; i += blockDim.x) neurons[i] = (double) a.codes[c * 64 + i] * 0.4;
    __syncthreads();
    seq_forward(a, neurons, thetas, true);                                                       // :229
    const double target = a.targets[c];
    if (threadIdx.x == 0) {
      const double om = __dadd_rn(target, -neurons[a.noff[out]]);                                // :242
      omegas[a.noff[out]] = om;
      s_err = __dmul_rn(om, om) / 2.0;                                                           // :243
    }
    __syncthreads();
    const double lambda = s_lambda;
    for (int layer = out - 1; layer >= 0; --layer) {                                             // :248
      const int nl = layer + 1, K = a.dims[layer], N = a.dims[nl];
      double *W = a.w + a.woff[layer], *dW = a.dw + a.woff[layer];
      // psi = omega * f'(theta); omega = 0  (:252-254). f'(theta) = 4*(1 - sig^2)/4 with sig = f(theta)/4 exactly.
      for (int j = threadIdx.x; j < N; j += blockDim.x) {
        psis[a.noff[nl] + j] = __dmul_rn(omegas[a.noff[nl] + j], act_df_from_a(neurons[a.noff[nl] + j]));
        omegas[a.noff[nl] + j] = 0.0;
      }
      __syncthreads();
      // omega[i] += psi_j * W[i][j] for j ascending, with the PRE-update weights (:257)
      const double *ps = psis + a.noff[nl];
      for (int i = threadIdx.x; i < K; i += blockDim.x) {
        double om = omegas[a.noff[layer] + i];
        const double *row = W + (long) i * N;
#pragma unroll 4
        for (int j = 0; j < N; ++j) om = __dadd_rn(om, __dmul_rn(ps[j], row[j]));
        omegas[a.noff[layer] + i] = om;
      }
      __syncthreads();
      // dW = (psi * lambda) * a ; W += dW  (:258-259)
      const double *in = neurons + a.noff[layer];
      const long total = (long) K * N;
      for (long e = threadIdx.x; e < total; e += blockDim.x) {
        const int i = (int) (e / N), j = (int) (e - (long) i * N);
        const double d = __dmul_rn(__dmul_rn(ps[j], lambda), in[i]);
        dW[e] = d;
        W[e] = __dadd_rn(W[e], d);
      }
      __syncthreads();
    }
    seq_forward(a, neurons, thetas, false);                                                      // :265
    if (threadIdx.x == 0) {
      const double d = __dadd_rn(target, -neurons[a.noff[out]]);
      s_new_err = __dmul_rn(d, d) / 2.0;                                                         // :266-267
      s_accept = s_new_err < s_err;
      if (s_accept) {
        if (s_lambda < a.max_const) s_lambda = __dmul_rn(s_lambda, a.rate);                      // :269-271
      } else {
        s_lambda = s_lambda / a.rate;                                                            // :273
      }
      if (a.accepted != nullptr) a.accepted[c] = (uint8_t) s_accept;
    }
    __syncthreads();
    if (!s_accept) {                                                                             // :274-277
      for (long e = threadIdx.x; e < a.n_weights; e += blockDim.x) a.w[e] = __dadd_rn(a.w[e], -a.dw[e]);
      __syncthreads();
    }
    if (a.apply_reset_rule) {                                                                    // train.cpp:54-57
      const bool reset = s_lambda < 0.0001 || 10.0 < s_lambda;
      __syncthreads();
      if (reset) {
        const unsigned long long offset = (unsigned long long) s_resets;
        for (long e = threadIdx.x; e < a.n_weights; e += blockDim.x) {                           // network.cpp:281-292
          uint32_t r[4];
          philox4x32_10((uint32_t) e, (uint32_t) ((uint64_t) e >> 32), (uint32_t) offset, (uint32_t) (offset >> 32),
                        (uint32_t) a.seed, (uint32_t) (a.seed >> 32), r);
          if (u53(r[0], r[1]) < 0.01) a.w[e] = 2.0 * u53(r[2], r[3]) + -1.0;
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_lambda = 2.0; ++s_resets; }
        __syncthreads();
      }
    }
    __threadfence();
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *a.lambda_io = s_lambda;
    if (a.resets != nullptr) *a.resets = s_resets;
    if (a.err_out != nullptr) { a.err_out[0] = s_err; a.err_out[1] = s_new_err; }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// minibatch: backward (psi / omega)
// ---------------------------------------------------------------------------------------------------------------
struct BackwardArgs {
  const uint8_t *pack;       // packed fragment stream of W^T (layers L-2 .. 1)
  const ChunkDesc *chunks;
  int n_chunks;
  int a_units;
  long n_boards;
  long n_groups;
  const double *targets;     // [n_boards]
  const double *acts;        // saved activations, unit layout
  double *psis;              // out: psi of every non-input layer, same geometry as acts
  long save_stride;
  uint32_t out_unit;         // unit offset of the output layer inside acts / psis
  int n_stages, stage_bytes;
  double *err_sum;           // += sum_b (T - y)^2 / 2
  const uint8_t *blob;       // the plan's chunk table, fetched by one bulk copy
  uint32_t fill_off[kMaxStages], fill_bytes[kMaxStages];   // the ring's first fills (no global read before the first fetch)
};

// psi = omega * f'(theta) on this warp's NBH tiles of a layer (network.cpp:252): omega sits in the pair buffer, the saved
// activation a = f(theta) comes from global memory. ALL NBH activation loads are issued before the first one is used (the
// accumulator registers are free at this point): one exposed L2 latency per layer boundary instead of one per tile
// (ncu of the per-tile loop: long_scoreboard 5.2 warps per issue, DMMA pipe 52 % active).
template <int NBH>
__device__ __forceinline__ void psi_tiles(double2 *a_out, const double *__restrict__ acts, double *__restrict__ psis, size_t goff0, size_t stride,
                                          int lane, bool valid) {
  double2 av[NBH];
#pragma unroll
  for (int j = 0; j < NBH; ++j) {
    av[j] = make_double2(0.0, 0.0);
    if (valid) av[j] = __ldcg(reinterpret_cast<const double2 *>(acts + goff0 + (size_t) j * stride));
  }
#pragma unroll
  for (int j = 0; j < NBH; ++j) {
    double2 om = a_out[j * 32 + lane];
    om.x = valid ? om.x * act_df_from_a(av[j].x) : 0.0;
    om.y = valid ? om.y * act_df_from_a(av[j].y) : 0.0;
    a_out[j * 32 + lane] = om;
    *reinterpret_cast<double2 *>(psis + goff0 + (size_t) j * stride) = om;
  }
}
#define CAB_NB_CASE(N) \
  case N: psi_tiles<N>(a_out, acts, psis, goff0, stride, lane, valid); break;
__device__ __forceinline__ void dispatch_psi(int nbh, double2 *a_out, const double *acts, double *psis, size_t goff0, size_t stride, int lane,
                                             bool valid) {
  switch (nbh) {
    CAB_NB_CASE(1) CAB_NB_CASE(2) CAB_NB_CASE(3) CAB_NB_CASE(4) CAB_NB_CASE(5) CAB_NB_CASE(6) CAB_NB_CASE(7)
    CAB_NB_CASE(8) CAB_NB_CASE(9) CAB_NB_CASE(10) CAB_NB_CASE(11) CAB_NB_CASE(12) CAB_NB_CASE(13) CAB_NB_CASE(14)
    CAB_NB_CASE(15) CAB_NB_CASE(16)
    default: break;
  }
}
#undef CAB_NB_CASE

__global__ void __launch_bounds__(kMaxThreads, 1) mlp_backward_kernel(const BackwardArgs args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int n_stages = args.n_stages, stage_bytes = args.stage_bytes;
  const int n_warps = (int) (blockDim.x >> 5), n_pairs = n_warps >> 1;
  uint8_t *ring = smem;
  double2 *abuf_all = reinterpret_cast<double2 *>(smem + fwd_smem_ring(n_stages, stage_bytes));
  ChunkDesc *s_chunks = reinterpret_cast<ChunkDesc *>(smem + fwd_smem_ring(n_stages, stage_bytes) + fwd_smem_abuf(n_pairs, args.a_units));
  // same carve-up as the forward kernel (its exp table and dot scratch lie between the chunk table and the barriers; unused here)
  uint64_t *full = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(s_chunks) + kBlobBytes + kDotScratchBytes);
  uint64_t *tbl_bar = full + n_stages;
  uint32_t *done = reinterpret_cast<uint32_t *>(tbl_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = warp >> 1, half = pair_half(warp);
  const int m = lane >> 2, q = lane & 3;
  const long my_groups = args.n_groups > blockIdx.x ? (args.n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long total_chunks = my_groups * args.n_chunks;
  const bool table_in_smem = args.n_chunks <= kSmemChunks;
  const ChunkDesc *chunks = table_in_smem ? s_chunks : args.chunks;
  // start-up trimmed like the forward kernel's (a CTA of an 8 192-board piece lives for one pass): the first group's
  // output and target are requested before anything else, the chunk table arrives by one bulk copy, the first ring fills
  // take their addresses from the kernel parameters
  double y_first = 0.0, t_first = 0.0;
  {
    const long board0 = ((long) blockIdx.x * n_pairs + pair) * kBoardsPerWarp;
    if (half == 0 && q == 0 && my_groups > 0 && board0 + m < args.n_boards) {
      y_first = __ldcg(args.acts + (size_t) args.out_unit * args.save_stride + (size_t) board0 * 8 + lane * 2);
      t_first = __ldg(args.targets + board0 + m);
    }
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) { mbar_init(&full[s], 1); done[s] = 0; }
    mbar_init(tbl_bar, 1);
    mbar_fence_init();
    for (int s = 0; s < n_stages && s < total_chunks; ++s) {
      mbar_arrive_expect_tx(&full[s], args.fill_bytes[s]);
      bulk_g2s(ring + (size_t) s * stage_bytes, args.pack + args.fill_off[s], args.fill_bytes[s], &full[s]);
    }
    if (table_in_smem) {
      mbar_arrive_expect_tx(tbl_bar, (uint32_t) kBlobChunkBytes);
      bulk_g2s(s_chunks, args.blob, (uint32_t) kBlobChunkBytes, tbl_bar);
    }
  }
  __syncthreads();

  double2 *abuf = abuf_all + (size_t) pair * args.a_units * 32;
  double acc[kMaxNBH][2];
#pragma unroll
  for (int j = 0; j < kMaxNBH; ++j) acc[j][0] = acc[j][1] = 0.0;
  uint32_t rs = 0, rk = 0;
  long n = 0;
  double err_local = 0.0;
  const uint32_t abuf_addr = smem_u32(abuf) + (uint32_t) lane * 16u;
  const uint32_t ring_addr = smem_u32(ring) + (uint32_t) lane * 16u;

  for (long g = blockIdx.x; g < args.n_groups; g += gridDim.x) {
    const long board0 = (g * n_pairs + pair) * kBoardsPerWarp;
    const long board = board0 + m;
    const bool valid = board < args.n_boards;
    const size_t lane_off = (size_t) board0 * 8 + lane * 2;

    // omega_out = T - y, E = omega^2/2 (network.cpp:242-243); psi_out = omega_out * f'(theta_out) (:252).
    // The output layer is one feature wide: it lives in .x of the lanes with q == 0 of unit 0.
    if (half == 0) {
      double2 psi = make_double2(0.0, 0.0);
      if (valid && q == 0) {
        const bool first_group = g == (long) blockIdx.x;
        const double y = first_group ? y_first : args.acts[(size_t) args.out_unit * args.save_stride + lane_off];
        const double om = (first_group ? t_first : args.targets[board]) - y;
        err_local += om * om / 2.0;
        psi.x = om * act_df_from_a(y);
      }
      abuf[lane] = psi;
      *reinterpret_cast<double2 *>(args.psis + (size_t) args.out_unit * args.save_stride + lane_off) = psi;
    }
    pair_barrier(pair);
    if (table_in_smem && g == (long) blockIdx.x) mbar_wait(tbl_bar, 0);   // the chunk table has landed

    for (int c = 0; c < args.n_chunks; ++c, ++n) {
      const ChunkDesc d = chunks[c];
      const bool last = d.flags & kLast;
      int j0, nbh;
      split_tiles(d.nb, d.layer, half, j0, nbh);
      mbar_wait(&full[rs], rk & 1);
      dispatch_chunk(nbh, acc, abuf_addr + (uint32_t) (d.a_in + d.t0) * kUnitBytes,
                     ring_addr + rs * (uint32_t) stage_bytes + (uint32_t) j0 * kUnitBytes, (uint32_t) d.nb * kUnitBytes, d.tcount);
      __syncwarp();
      if (lane == 0) {
        const uint32_t old = atomicAdd(&done[rs], 1u);
        if ((old + 1) % (uint32_t) n_warps == 0) {
          const long refill = n + n_stages;
          if (refill < total_chunks) {
            const ChunkDesc *nd = &chunks[refill % args.n_chunks];
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive_expect_tx(&full[rs], nd->bytes);
            bulk_g2s(ring + (size_t) rs * stage_bytes, args.pack + nd->src_off, nd->bytes, &full[rs]);
          }
        }
      }
      if (++rs == (uint32_t) n_stages) { rs = 0; ++rk; }

      if (last) {
        // the accumulators hold omega of layer d.layer (network.cpp:257); psi = omega * f'(theta) uses the saved
        // activation a = f(theta) of that layer (:252). psi goes to the pair buffer (next contraction) and to global.
        double2 *a_out = abuf + (size_t) (d.a_out + j0) * 32;
        pair_barrier(pair);
        dispatch_store(nbh, acc, a_out, lane);
#pragma unroll
        for (int j = 0; j < kMaxNBH; ++j) acc[j][0] = acc[j][1] = 0.0;   // the next chunk starts a layer: zeroed HERE, so that the
                                                                          // registers are visibly free for the loads below
        dispatch_psi(nbh, a_out, args.acts, args.psis, (size_t) (d.save_off + d.tile0 + j0) * args.save_stride + lane_off, (size_t) args.save_stride,
                     lane, valid);
        pair_barrier(pair);
      }
    }
  }
  if (args.err_sum != nullptr) {
    for (int o = 16; o > 0; o >>= 1) err_local += __shfl_xor_sync(0xffffffffu, err_local, o);
    if (lane == 0 && err_local != 0.0) atomicAdd(args.err_sum, err_local);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// minibatch: weight gradient  G_l[k][n] += sum_b a_l[b][k] * psi_{l+1}[b][n]
// One warp owns a block of up to kGK x kGN 8x8 tiles of one layer and a slice of the boards (an "item"); per step of
// 4 boards it needs kGK A fragments + kGN B fragments (256 contiguous bytes each: the unit layout is
// [unit][board][8]) and issues kGK*kGN DMMAs (M = k features, N = n features, K = boards).
// The fragments reach the warp through its PRIVATE shared-memory ring: kGradStages stages of kGradStageBoards boards,
// each filled by one 1-D TMA bulk copy per unit (1 KB) completing on the stage's mbarrier. With plain global loads
// the kernel sat on long_scoreboard (ncu: 5.0 of 8.4 stalled warps per issue, DMMA pipe 52 % active on 8 warps/SM).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kGK = 4, kGN = 8;
constexpr int kGradWarps = 8;               // per CTA (default; GradArgs::n_warps), one CTA per SM
constexpr int kGradStageBoards = 16;        // 4 DMMA steps
constexpr int kGradStages = 2;              // default; GradArgs::n_stages
constexpr int kGradUnitBytes = kGradStageBoards * 64;                       // one unit of one stage
constexpr int kGradStageBytes = (kGK + kGN) * kGradUnitBytes;               // 12 KB
constexpr size_t kGradSmem = (size_t) kGradWarps * kGradStages * kGradStageBytes + kGradWarps * kGradStages * sizeof(uint64_t);
struct GradTask {
  uint32_t a_unit, p_unit;   // first unit of the k-tiles inside acts, of the n-tiles inside psis
  int nkt, nnt;              // tiles in this block (<= kGK, <= kGN)
  int k0, n0;                // first feature indices
  int K, N;                  // layer shape
  long w_off;
};
struct GradItem {            // one warp's work: a task and a board range (multiple of 4 boards)
  int task;
  int b_lo, b_hi;
  int pad;
};
struct GradArgs {
  const GradTask *tasks;
  const GradItem *items;
  int n_items;
  int n_warps, n_stages;
  int *next_item;            // global work counter, zeroed before the launch
  const double *acts, *psis;
  long save_stride;
  double *grad;
};

// One item with a compile-time block shape: no predicates around the DMMAs (the runtime-shaped loop spent 4-5
// instructions per DMMA on predicate logic and issued at 22 cycles per DMMA instead of 16).
template <int NKT, int NNT>
__device__ __forceinline__ void grad_item(const GradArgs &args, const GradItem &it, const GradTask &t, uint8_t *ring, uint64_t *full,
                                          uint32_t &parity_bits, int ring_units, int lane) {
  constexpr int n_units = NKT + NNT;
  constexpr uint32_t stage_bytes = (uint32_t) n_units * kGradUnitBytes;   // stages packed tightly: thin blocks get a deeper ring
  const int item_stages = min(ring_units / n_units, 8);
  // lane u < n_units copies unit u of a stage (A units first, then B units), 1 KB each
  const double *my_src = nullptr;
  if (lane < n_units)
    my_src = lane < NKT ? args.acts + (size_t) (t.a_unit + lane) * args.save_stride
                        : args.psis + (size_t) (t.p_unit + (lane - NKT)) * args.save_stride;
  auto fill = [&](int stage, long board0) {
    if (lane == 0) mbar_arrive_expect_tx(&full[stage], stage_bytes);
    __syncwarp();
    if (lane < n_units)
      bulk_g2s(ring + (size_t) stage * stage_bytes + (size_t) lane * kGradUnitBytes, my_src + (size_t) board0 * 8, kGradUnitBytes, &full[stage]);
  };
  const int n_stage_iters = (it.b_hi - it.b_lo + kGradStageBoards - 1) / kGradStageBoards;
  for (int s = 0; s < item_stages && s < n_stage_iters; ++s) fill(s, it.b_lo + (long) s * kGradStageBoards);

  double acc[NKT][NNT][2];
#pragma unroll
  for (int i = 0; i < NKT; ++i)
#pragma unroll
    for (int j = 0; j < NNT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  // lane (r = lane>>2, c = lane&3): A[k = 8kt + r][board b + c], B[board b + c][n = 8nt + r]
  const uint32_t lane_off = (uint32_t) ((lane & 3) * 64 + (lane >> 2) * 8);
  int stage = 0;
  for (int si = 0; si < n_stage_iters; ++si) {
    const uint8_t *base = ring + (size_t) stage * stage_bytes + lane_off;
    mbar_wait(&full[stage], (parity_bits >> stage) & 1u);
    parity_bits ^= 1u << stage;
    const int steps = min(kGradStageBoards / 4, (it.b_hi - (it.b_lo + si * kGradStageBoards)) / 4);
    double af[NKT], bf[NNT], an[NKT], bn[NNT];
#pragma unroll
    for (int i = 0; i < NKT; ++i) af[i] = *reinterpret_cast<const double *>(base + i * kGradUnitBytes);
#pragma unroll
    for (int j = 0; j < NNT; ++j) bf[j] = *reinterpret_cast<const double *>(base + (NKT + j) * kGradUnitBytes);
    for (int st = 0; st < steps; ++st) {
      const int nx = st + 1 < steps ? st + 1 : st;
#pragma unroll
      for (int i = 0; i < NKT; ++i) an[i] = *reinterpret_cast<const double *>(base + i * kGradUnitBytes + nx * 256);
#pragma unroll
      for (int j = 0; j < NNT; ++j) bn[j] = *reinterpret_cast<const double *>(base + (NKT + j) * kGradUnitBytes + nx * 256);
#pragma unroll
      for (int i = 0; i < NKT; ++i)
#pragma unroll
        for (int j = 0; j < NNT; ++j) dmma884(acc[i][j], af[i], bf[j]);
#pragma unroll
      for (int i = 0; i < NKT; ++i) af[i] = an[i];
#pragma unroll
      for (int j = 0; j < NNT; ++j) bf[j] = bn[j];
    }
    __syncwarp();                               // every lane has its fragments in registers: the stage may be refilled
    if (si + item_stages < n_stage_iters) fill(stage, it.b_lo + (long) (si + item_stages) * kGradStageBoards);
    if (++stage == item_stages) stage = 0;
  }
  // C[row = lane>>2][col = 2*(lane&3) + {0,1}] -> G[k][n]
#pragma unroll
  for (int i = 0; i < NKT; ++i)
#pragma unroll
    for (int j = 0; j < NNT; ++j) {
      const int k = t.k0 + 8 * i + (lane >> 2), nn = t.n0 + 8 * j + 2 * (lane & 3);
      if (k < t.K) {
        if (nn < t.N) atomicAdd(args.grad + t.w_off + (long) k * t.N + nn, acc[i][j][0]);
        if (nn + 1 < t.N) atomicAdd(args.grad + t.w_off + (long) k * t.N + nn + 1, acc[i][j][1]);
      }
    }
}

__global__ void __launch_bounds__(kGradWarps * 32, 1) grad_kernel(const GradArgs args) {
  extern __shared__ __align__(128) uint8_t grad_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_warps = args.n_warps, n_stages = args.n_stages;
  uint8_t *ring = grad_smem + (size_t) warp * n_stages * kGradStageBytes;
  uint64_t *full = reinterpret_cast<uint64_t *>(grad_smem + (size_t) n_warps * n_stages * kGradStageBytes) + warp * 8;
  if (lane == 0) {
    for (int s = 0; s < 8; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncwarp();
  const int ring_units = n_stages * (kGK + kGN);   // ring capacity in 1 KB unit slots
  uint32_t parity_bits = 0;                        // per-stage phase parity of this warp's mbarriers
  // persistent warps: each one draws items from a global counter until none are left (no CTA-wide barrier anywhere)
  for (;;) {
    int item_idx = 0;
    if (lane == 0) item_idx = atomicAdd(args.next_item, 1);
    item_idx = __shfl_sync(0xffffffffu, item_idx, 0);
    if (item_idx >= args.n_items) break;
    const GradItem it = args.items[item_idx];
    const GradTask t = args.tasks[it.task];
#define CAB_GRAD_CASE(A, B) case (A) * 16 + (B): grad_item<A, B>(args, it, t, ring, full, parity_bits, ring_units, lane); break;
#define CAB_GRAD_ROW(A) CAB_GRAD_CASE(A, 1) CAB_GRAD_CASE(A, 2) CAB_GRAD_CASE(A, 3) CAB_GRAD_CASE(A, 4) \
                        CAB_GRAD_CASE(A, 5) CAB_GRAD_CASE(A, 6) CAB_GRAD_CASE(A, 7) CAB_GRAD_CASE(A, 8)
    switch (t.nkt * 16 + t.nnt) {
      CAB_GRAD_ROW(1) CAB_GRAD_ROW(2) CAB_GRAD_ROW(3) CAB_GRAD_ROW(4)
      default: break;
    }
#undef CAB_GRAD_ROW
#undef CAB_GRAD_CASE
  }
}

// dW = (lambda / count) * G ; W += dW     (batch mean of network.cpp:258-259). count = grad[n + 2].
__global__ void __launch_bounds__(256) apply_update_kernel(double *__restrict__ w, double *__restrict__ dw,
                                                            const double *__restrict__ grad, long n, double lambda) {
  const double scale = lambda / grad[n + 2];
  for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
    const double d = grad[i] * scale;
    dw[i] = d;
    w[i] += d;
  }
}
// ---- the minibatch step's decision, on the device (no host round trip inside a step) -----------------------------------
struct TrainCtl {              // device-resident state of the minibatch trainer
  double lambda, max_const, rate;
  double e0, e1;               // batch-mean E and E' of the last step
  double e_first;              // E of the first step since the state was reset
  int ok;                      // last step accepted
  int reset_rule;              // train.cpp:54-57 after every step: lambda outside [1e-4, 10] -> lambda = 2, dropout
  int do_dropout;              // set by decide for the dropout kernel that follows
  int resets;
  long long accepted, steps;
  unsigned long long seed;     // Philox key of the dropout
  int peer_error;              // a peer's flag did not arrive in time (cab_dp_peer_*): the results are not valid
};
// dW = (lambda / count) * G ; W += dW with lambda read on the device
__global__ void __launch_bounds__(256) apply_update_ctl_kernel(double *__restrict__ w, double *__restrict__ dw, const double *__restrict__ grad,
                                                                long n, const TrainCtl *__restrict__ ctl) {
  const double scale = ctl->lambda / grad[n + 2];
  for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
    const double d = grad[i] * scale;
    dw[i] = d;
    w[i] += d;
  }
}
// E' < E ? accept (lambda *= rate while lambda < max_const) : reject (lambda /= rate; the rollback kernel undoes the update)
// (network.cpp:269-278), then the train loop's reset rule (train.cpp:54-57). grad_tail = {sum E, -, count}; e1_sum = sum E'.
__device__ __forceinline__ void decide_body(TrainCtl *ctl, double e0_sum, double count, double e1_sum, double *e_trace) {
  const double e0 = e0_sum / count, e1 = e1_sum / count;
  const bool ok = e1 < e0;
  double lam = ctl->lambda;
  if (ok) { if (lam < ctl->max_const) lam *= ctl->rate; }
  else lam /= ctl->rate;
  ctl->do_dropout = 0;
  if (ctl->reset_rule && (lam < 0.0001 || 10.0 < lam)) {
    lam = 2.0;
    ctl->do_dropout = 1;
    ctl->resets += 1;
  }
  if (e_trace != nullptr) { e_trace[2 * ctl->steps] = e0; e_trace[2 * ctl->steps + 1] = e1; }
  if (ctl->steps == 0) ctl->e_first = e0;
  ctl->lambda = lam;
  ctl->e0 = e0; ctl->e1 = e1;
  ctl->ok = ok ? 1 : 0;
  ctl->accepted += ok ? 1 : 0;
  ctl->steps += 1;
}
__global__ void decide_kernel(TrainCtl *ctl, const double *grad_tail, const double *e1_sum, double *e_trace) {
  decide_body(ctl, grad_tail[0], grad_tail[2], *e1_sum, e_trace);
}

// ---- data-parallel step over NVLink peer memory: the collective fused into the update --------------------------------
// Every rank owns an exchange block {grad[2][nw + 4], e1[2][kPeerMax], flags_grad[kPeerMax], flags_e1[kPeerMax]} that
// all ranks have mapped (CUDA IPC). Step e (epoch, from 1) uses the buffers of parity e & 1:
//   peer_signal_kernel         after the gradient kernel: "my grad[e & 1] is complete" = e into flags_grad[me] of every rank
//   peer_reduce_update_kernel  waits for all ranks' flags, then every rank sums the ranks' gradients itself, element by
//                              element IN RANK ORDER (identical bits on every rank), and applies dW = (lambda / count) sum:
//                              the all-reduce and apply_update_ctl_kernel in one pass over the weights, remote reads only
//   peer_decide_kernel         E' the same way: 8 bytes stored into every rank's e1[e & 1][me] + flag, then decide_kernel
// A buffer of parity p is rewritten two steps later; by then every rank has passed the next step's flag wait, which each
// rank only signals after it has finished reading (stream order): one flag per step and buffer suffices.
// Waits give up after ~2 s (a dead peer must not hang the GPU): TrainCtl::peer_error, reported by the host.
constexpr int kPeerMax = 16;
struct PeerArgs {
  void *const *table;                        // [world] every rank's exchange block in this rank's address space
  int rank, world;
  unsigned long long epoch;
  size_t off_grad[2], off_e1, off_flags_grad, off_flags_e1;
  long nw;
};
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_peer_f64(const double *p) {   // never from a stale L1 line
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
// lanes r < world of the calling warp wait until flags[r] >= epoch; false (warp-uniform) on timeout
__device__ __forceinline__ bool peer_wait(const unsigned long long *flags, int world, unsigned long long epoch) {
  const int r = threadIdx.x & 31;
  bool ok = true;
  if (r < world) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flags + r) < epoch)
      if (clock64() - t0 > 4000000000ll) { ok = false; break; }
  }
  return __all_sync(0xffffffffu, ok);
}
__global__ void peer_signal_kernel(const PeerArgs a) {             // <<<1, 32>>>
  __threadfence_system();                                           // the gradient kernel's results, system-wide, before the flag
  const int r = threadIdx.x;
  if (r < a.world)
    st_release_sys(reinterpret_cast<unsigned long long *>(static_cast<char *>(a.table[r]) + a.off_flags_grad) + a.rank, a.epoch);
}
__global__ void __launch_bounds__(256) peer_reduce_update_kernel(const PeerArgs a, double *__restrict__ w, double *__restrict__ dw, TrainCtl *ctl,
                                                                  double *__restrict__ tail_out) {
  __shared__ int s_ok;
  __shared__ double s_tail[4];
  const int p = (int) (a.epoch & 1);
  if (threadIdx.x < 32) {
    const bool ok = peer_wait(reinterpret_cast<const unsigned long long *>(static_cast<const char *>(a.table[a.rank]) + a.off_flags_grad), a.world, a.epoch);
    if (threadIdx.x == 0) s_ok = ok ? 1 : 0;
  }
  __syncthreads();
  if (!s_ok) {
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->peer_error = 1;
    return;
  }
  if (threadIdx.x < 4) {                      // sum E, -, count, spare over the ranks
    double t = 0.0;
    for (int r = 0; r < a.world; ++r) t += ld_peer_f64(reinterpret_cast<const double *>(static_cast<const char *>(a.table[r]) + a.off_grad[p]) + a.nw + threadIdx.x);
    s_tail[threadIdx.x] = t;
    if (blockIdx.x == 0) tail_out[threadIdx.x] = t;
  }
  __syncthreads();
  const double scale = ctl->lambda / s_tail[2];
  const double *g[kPeerMax];
#pragma unroll
  for (int r = 0; r < kPeerMax; ++r) g[r] = r < a.world ? reinterpret_cast<const double *>(static_cast<const char *>(a.table[r]) + a.off_grad[p]) : nullptr;
  for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < a.nw; i += (long) gridDim.x * blockDim.x) {
    double sum = 0.0;
#pragma unroll
    for (int r = 0; r < kPeerMax; ++r)
      if (r < a.world) sum += ld_peer_f64(g[r] + i);
    const double d = sum * scale;
    dw[i] = d;
    w[i] += d;
  }
}
__global__ void peer_decide_kernel(const PeerArgs a, TrainCtl *ctl, const double *tail, const double *e1_local, double *e_trace) {   // <<<1, 32>>>
  const int r = threadIdx.x, p = (int) (a.epoch & 1);
  if (r < a.world) {
    double *slot = reinterpret_cast<double *>(static_cast<char *>(a.table[r]) + a.off_e1) + p * kPeerMax + a.rank;
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(slot), "d"(*e1_local) : "memory");
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned long long *>(static_cast<char *>(a.table[r]) + a.off_flags_e1) + a.rank, a.epoch);
  }
  const char *mine = static_cast<const char *>(a.table[a.rank]);
  const bool ok = peer_wait(reinterpret_cast<const unsigned long long *>(mine + a.off_flags_e1), a.world, a.epoch);
  if (threadIdx.x != 0) return;
  if (!ok || ctl->peer_error) { ctl->peer_error = 1; return; }
  double e1 = 0.0;
  for (int q = 0; q < a.world; ++q) e1 += ld_peer_f64(reinterpret_cast<const double *>(mine + a.off_e1) + p * kPeerMax + q);
  decide_body(ctl, tail[0], tail[2], e1, e_trace);
}
__global__ void __launch_bounds__(256) rollback_ctl_kernel(double *__restrict__ w, const double *__restrict__ dw, long n,
                                                            const TrainCtl *__restrict__ ctl) {
  if (ctl->ok) return;
  for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) w[i] -= dw[i];
}
// Optimizer::dropout behind the reset rule: weight i is replaced w.p. 0.01 by U(-1,1), Philox(seed, offset = reset index)
__global__ void __launch_bounds__(256) dropout_ctl_kernel(double *__restrict__ w, long n, double rate, const TrainCtl *__restrict__ ctl) {
  if (!ctl->do_dropout) return;
  const unsigned long long offset = (unsigned long long) (ctl->resets - 1), seed = ctl->seed;
  for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
    uint32_t r[4];
    philox4x32_10((uint32_t) i, (uint32_t) ((uint64_t) i >> 32), (uint32_t) offset, (uint32_t) (offset >> 32), (uint32_t) seed, (uint32_t) (seed >> 32), r);
    if (u53(r[0], r[1]) < rate) w[i] = 2.0 * u53(r[2], r[3]) + -1.0;
  }
}

// W -= dW   (network.cpp:274-277)
__global__ void __launch_bounds__(256) rollback_kernel(double *__restrict__ w, const double *__restrict__ dw, long n) {
  for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) w[i] -= dw[i];
}
// sum_b (T - y)^2 / 2 over a batch of network outputs
__global__ void __launch_bounds__(256) batch_error_kernel(const double *__restrict__ y, const double *__restrict__ t, long n,
                                                           double *__restrict__ err_sum) {
  double e = 0.0;
  for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
    const double d = t[i] - y[i];
    e += d * d / 2.0;
  }
  for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
  if ((threadIdx.x & 31) == 0 && e != 0.0) atomicAdd(err_sum, e);
}
__global__ void set_scalar_kernel(double *p, double v) { *p = v; }

}  // namespace cab
