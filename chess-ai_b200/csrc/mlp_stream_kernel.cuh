// mlp_stream_kernel.cuh — the fused whole-MLP forward kernel (K1 / K2).
//
// Replaces Network::predictPosition / propagate_for_training (/root/reference/src/mcts_network/network.cpp:92-153):
//   per layer, per output j:  sum_i a[i] * W[i][j]  ->  f(sum)
// B200 design:
//   * a PAIR of warps owns 8 boards through ALL layers; each warp of the pair computes half of every layer's output
//     tiles. Halving the accumulators keeps a thread under 128 registers, so 16 warps (4 per SMSP) are resident and
//     one warp's loads / barriers / activation math hide behind the others' DMMAs. The pair shares one activation
//     buffer in shared memory (8 boards x widest layer) and meets at a 64-thread named barrier per layer boundary
//   * the matmuls run on the fp64 tensor pipe: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4), accumulators in registers
//   * weights arrive as a pre-packed B-fragment stream, moved L2 -> shared by 1-D TMA bulk copies
//     (cp.async.bulk, SASS UBLKCP) into an mbarrier ring shared by the CTA. No producer warp: the LAST warp to
//     finish a chunk (shared-memory arrival counter) re-arms that ring slot on the spot
//   * the DMMA C-fragment of tile j IS the A-fragment pair of the next layer's k-steps (2j, 2j+1) under the feature
//     permutation k = 8t + 2q + h that the packed stream bakes in, so activations never change layout
//   * persistent grid: CTAs loop over board groups; the weight stream is re-read from L2 per group
//   * a CTA of a 4096-board launch lives for one pass: its start-up (input codes requested first, tables by one bulk
//     copy, first ring fills from kernel parameters) is trimmed like an inner loop
#pragma once
#include "cab_internal.cuh"

namespace cab {

// shared-memory fragment load with a pinned position in the instruction stream
__device__ __forceinline__ double2 lds_frag(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void dmma_v(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void pair_barrier(int pair) {  // named barrier 1 + pair, 64 threads (SASS BAR.SYNC)
  asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory");
}

// One chunk = tcount input units x NBH output tiles of this warp's half. B fragments are fetched one group of kG
// tiles ahead (across the t boundary too); inside a group the kG ".x" DMMAs are issued before the dependent ".y" ones.
constexpr int kG = 2;
template <int NBH>
__device__ __forceinline__ void run_chunk(double (&acc)[kMaxNBH][2], uint32_t a_addr, uint32_t w_addr, uint32_t t_stride, int tcount) {
  constexpr int NG = (NBH + kG - 1) / kG;
  double2 bcur[kG], bnxt[kG];
  double2 a = lds_frag(a_addr);
#pragma unroll
  for (int i = 0; i < kG; ++i)
    if (i < NBH) bcur[i] = lds_frag(w_addr + i * kUnitBytes);
  for (int t = 0; t < tcount; ++t) {
    const uint32_t wt = w_addr + (uint32_t) t * t_stride;
    const bool more = t + 1 < tcount;
    double2 a_next = a;
    if (more) a_next = lds_frag(a_addr + (uint32_t) (t + 1) * kUnitBytes);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      if (g + 1 < NG) {
#pragma unroll
        for (int i = 0; i < kG; ++i)
          if ((g + 1) * kG + i < NBH) bnxt[i] = lds_frag(wt + ((g + 1) * kG + i) * kUnitBytes);
      } else if (more) {
#pragma unroll
        for (int i = 0; i < kG; ++i)
          if (i < NBH) bnxt[i] = lds_frag(wt + t_stride + i * kUnitBytes);
      }
#pragma unroll
      for (int i = 0; i < kG; ++i)
        if (g * kG + i < NBH) dmma_v(acc[g * kG + i], a.x, bcur[i].x);
#pragma unroll
      for (int i = 0; i < kG; ++i)
        if (g * kG + i < NBH) dmma_v(acc[g * kG + i], a.y, bcur[i].y);
#pragma unroll
      for (int i = 0; i < kG; ++i) bcur[i] = bnxt[i];
    }
    a = a_next;
  }
}

template <int NBH>
__device__ __forceinline__ void store_acc(const double (&acc)[kMaxNBH][2], double2 *__restrict__ a_out, int lane) {
#pragma unroll
  for (int j = 0; j < NBH; ++j) a_out[j * 32 + lane] = make_double2(acc[j][0], acc[j][1]);
}

#define CAB_NB_CASE(N) \
  case N: run_chunk<N>(acc, a_addr, w_addr, t_stride, tcount); break;
__device__ __forceinline__ void dispatch_chunk(int nbh, double (&acc)[kMaxNBH][2], uint32_t a_addr, uint32_t w_addr,
                                               uint32_t t_stride, int tcount) {
  switch (nbh) {
    CAB_NB_CASE(1) CAB_NB_CASE(2) CAB_NB_CASE(3) CAB_NB_CASE(4) CAB_NB_CASE(5) CAB_NB_CASE(6) CAB_NB_CASE(7)
    CAB_NB_CASE(8) CAB_NB_CASE(9) CAB_NB_CASE(10) CAB_NB_CASE(11) CAB_NB_CASE(12) CAB_NB_CASE(13) CAB_NB_CASE(14)
    CAB_NB_CASE(15) CAB_NB_CASE(16)
    default: break;
  }
}
#undef CAB_NB_CASE
// f() on accumulator tiles T0 .. T0+NT-1 in place (2 NT independent activations in flight per lane)
template <int T0, int NT>
__device__ __forceinline__ void act_tiles(double (&acc)[kMaxNBH][2], const double *tbl) {
#pragma unroll
  for (int u = 0; u < NT; ++u) {
    acc[T0 + u][0] = act_f_tbl(acc[T0 + u][0], tbl);
    acc[T0 + u][1] = act_f_tbl(acc[T0 + u][1], tbl);
  }
}
// f() on this warp's nbh tiles, four tiles (eight activations in flight) at a time and EXACTLY nbh of them: the
// activation shares the fp64 pipe with the DMMAs, a padded tile costs as much as a real one. The switch gives the
// rolled loop static register indices.
__device__ __forceinline__ void act_all(int nbh, double (&acc)[kMaxNBH][2], const double *tbl) {
  for (int t0 = 0; t0 < nbh; t0 += 4) {
    const int nt = nbh - t0 < 4 ? nbh - t0 : 4;
    switch (t0 + nt - 1) {
      case 0: act_tiles<0, 1>(acc, tbl); break;
      case 1: act_tiles<0, 2>(acc, tbl); break;
      case 2: act_tiles<0, 3>(acc, tbl); break;
      case 3: act_tiles<0, 4>(acc, tbl); break;
      case 4: act_tiles<4, 1>(acc, tbl); break;
      case 5: act_tiles<4, 2>(acc, tbl); break;
      case 6: act_tiles<4, 3>(acc, tbl); break;
      case 7: act_tiles<4, 4>(acc, tbl); break;
      case 8: act_tiles<8, 1>(acc, tbl); break;
      case 9: act_tiles<8, 2>(acc, tbl); break;
      case 10: act_tiles<8, 3>(acc, tbl); break;
      case 11: act_tiles<8, 4>(acc, tbl); break;
      case 12: act_tiles<12, 1>(acc, tbl); break;
      case 13: act_tiles<12, 2>(acc, tbl); break;
      case 14: act_tiles<12, 3>(acc, tbl); break;
      default: act_tiles<12, 4>(acc, tbl); break;
    }
  }
}

// pre-activations theta of this warp's tiles to global memory, unit layout (K2 with theta export, network.cpp:139-153)
template <int NBH>
__device__ __forceinline__ void store_acc_global(const double (&acc)[kMaxNBH][2], double *__restrict__ dst, long stride) {
#pragma unroll
  for (int j = 0; j < NBH; ++j) *reinterpret_cast<double2 *>(dst + (size_t) j * stride) = make_double2(acc[j][0], acc[j][1]);
}
#define CAB_NB_CASE(N) \
  case N: store_acc_global<N>(acc, dst, stride); break;
__device__ __forceinline__ void dispatch_store_global(int nbh, const double (&acc)[kMaxNBH][2], double *dst, long stride) {
  switch (nbh) {
    CAB_NB_CASE(1) CAB_NB_CASE(2) CAB_NB_CASE(3) CAB_NB_CASE(4) CAB_NB_CASE(5) CAB_NB_CASE(6) CAB_NB_CASE(7)
    CAB_NB_CASE(8) CAB_NB_CASE(9) CAB_NB_CASE(10) CAB_NB_CASE(11) CAB_NB_CASE(12) CAB_NB_CASE(13) CAB_NB_CASE(14)
    CAB_NB_CASE(15) CAB_NB_CASE(16)
    default: break;
  }
}
#undef CAB_NB_CASE

// sum_j acc[j] . w[j] for this warp's NBH tiles; w_addr = this lane's 16 bytes of the first tile's unit
template <int NBH>
__device__ __forceinline__ double dot_acc(const double (&acc)[kMaxNBH][2], uint32_t w_addr) {
  double s0 = 0.0, s1 = 0.0;
#pragma unroll
  for (int j = 0; j < NBH; ++j) {
    const double2 w = lds_frag(w_addr + j * kUnitBytes);
    s0 = fma(acc[j][0], w.x, s0);
    s1 = fma(acc[j][1], w.y, s1);
  }
  return s0 + s1;
}
#define CAB_NB_CASE(N) \
  case N: return dot_acc<N>(acc, w_addr);
__device__ __forceinline__ double dispatch_dot(int nbh, const double (&acc)[kMaxNBH][2], uint32_t w_addr) {
  switch (nbh) {
    CAB_NB_CASE(1) CAB_NB_CASE(2) CAB_NB_CASE(3) CAB_NB_CASE(4) CAB_NB_CASE(5) CAB_NB_CASE(6) CAB_NB_CASE(7)
    CAB_NB_CASE(8) CAB_NB_CASE(9) CAB_NB_CASE(10) CAB_NB_CASE(11) CAB_NB_CASE(12) CAB_NB_CASE(13) CAB_NB_CASE(14)
    CAB_NB_CASE(15) CAB_NB_CASE(16)
    default: return 0.0;
  }
}
#undef CAB_NB_CASE

#define CAB_NB_CASE(N) \
  case N: store_acc<N>(acc, a_out, lane); break;
__device__ __forceinline__ void dispatch_store(int nbh, const double (&acc)[kMaxNBH][2], double2 *a_out, int lane) {
  switch (nbh) {
    CAB_NB_CASE(1) CAB_NB_CASE(2) CAB_NB_CASE(3) CAB_NB_CASE(4) CAB_NB_CASE(5) CAB_NB_CASE(6) CAB_NB_CASE(7)
    CAB_NB_CASE(8) CAB_NB_CASE(9) CAB_NB_CASE(10) CAB_NB_CASE(11) CAB_NB_CASE(12) CAB_NB_CASE(13) CAB_NB_CASE(14)
    CAB_NB_CASE(15) CAB_NB_CASE(16)
    default: break;
  }
}
#undef CAB_NB_CASE

// Which warp of a pair is "half 0", and how a block of nb output tiles is split between the halves.
//  * halves are interleaved over the SM sub-partitions (warp w sits on SMSP w % 4): with half = w & 1, SMSPs 0 and 2
//    would host only half-0 warps and carry every layer's odd tile
//  * the odd tile of a layer goes to half (layer & 1), so the halves' totals balance over the network
//    (default net: 527 vs 529 tile-units per pass instead of 558 vs 498)
__device__ __forceinline__ int pair_half(int warp) { return (warp ^ (warp >> 2)) & 1; }
__device__ __forceinline__ void split_tiles(int nb, int layer, int half, int &j0, int &nbh) {
  const int small = nb >> 1, big = nb - small;
  const int share0 = (layer & 1) == 0 ? big : small;   // tiles of half 0 (first in the block)
  j0 = half ? share0 : 0;
  nbh = half ? nb - share0 : share0;
}

constexpr int kSmemChunks = 96;  // chunk descriptors cached in shared memory (default net: 21); longer plans read L2

struct ForwardArgs {
  const uint8_t *pack;      // packed fragment stream of W
  const ChunkDesc *chunks;
  int n_chunks;
  int a_units;              // per-pair activation buffer, in units
  long n_boards;
  long n_groups;            // ceil(n_boards / (8 * pairs per CTA))
  const int8_t *codes;      // [n_boards][64], 2-byte aligned
  double *out;              // [n_boards] or nullptr
  double *save;             // saved activations in unit layout, or nullptr (K2: propagate_for_training)
  double *save_theta;       // saved PRE-activations theta (network.cpp:150 unbounded[n][j] = sum), same layout; THETA kernels only
  long save_stride;         // doubles per unit row = boards_pad * 8
  int n_stages;             // weight ring depth
  int stage_bytes;          // bytes per ring slot (the plan's chunk size limit)
  long long *prof;          // optional [n_warps_total][8] phase cycle counters (CAB_FWD_PROFILE debug launches)
  int team_shift;           // warps >> team_shift = team; team k starts k * team_lag cycles late (0 lag: lock-step)
  int team_lag;
  const uint8_t *blob;      // [ChunkDesc x kSmemChunks][exp table]: the kernel's tables, fetched by ONE bulk copy
  uint32_t fill_off[kMaxStages], fill_bytes[kMaxStages];   // the first n_stages chunks (kernel parameters: no global read before the first fetch)
};

// shared memory carve-up, in this order
__host__ __device__ inline size_t fwd_smem_ring(int n_stages, int stage_bytes) { return (size_t) n_stages * stage_bytes; }
__host__ __device__ inline size_t fwd_smem_abuf(int n_pairs, int a_units) { return (size_t) n_pairs * a_units * kUnitBytes; }
constexpr size_t kBlobChunkBytes = (size_t) kSmemChunks * sizeof(ChunkDesc), kBlobBytes = kBlobChunkBytes + kExpTblSize * sizeof(double);
constexpr size_t kDotScratchBytes = (size_t) (kConsumerWarps / 2) * kBoardsPerWarp * sizeof(double);   // pair exchange of the fused output layer
__host__ __device__ inline size_t fwd_smem_tail(int n_stages) {
  return kBlobBytes + kDotScratchBytes + (size_t) (n_stages + 1) * sizeof(uint64_t) + (size_t) n_stages * sizeof(uint32_t) + 16;
}

#define CAB_PROF_T(var) long long var = 0; if (PROF) var = clock64();
#define CAB_PROF_ACC(slot, t0) if (PROF) { prof_acc[slot] += clock64() - (t0); }

// the input codes of one lane for one board group: 4 x 2 bytes (units 2 tt + half, features 2q, 2q+1 of board m)
struct LaneCodes { uint32_t v[2]; };   // bytes: unit half+0: (c0, c1), half+2: (c0, c1) | half+4, half+6
__device__ __forceinline__ LaneCodes load_codes(const int8_t *codes, long board, bool valid, int half, int q) {
  LaneCodes c;
  const int8_t *row = codes + (valid ? board : 0) * 64;
  uint32_t w[4];
#pragma unroll
  for (int tt = 0; tt < 4; ++tt) {
    w[tt] = 0;
    if (valid) w[tt] = __ldg(reinterpret_cast<const uint16_t *>(row + 8 * (2 * tt + half) + 2 * q));
  }
  c.v[0] = w[0] | (w[1] << 16);
  c.v[1] = w[2] | (w[3] << 16);
  return c;
}

// SAVE = true: K2, every layer's activations also go to args.save (propagate_for_training, network.cpp:139-153).
// THETA = true (with SAVE): the pre-activations too, to args.save_theta (K2 as the reference's API returns it).
// The plain evaluator (SAVE = false) carries none of that code or its registers.
template <bool PROF, bool SAVE = false, bool THETA = false>
__global__ void __launch_bounds__(kMaxThreads, 1) mlp_forward_kernel(const ForwardArgs args) {
  long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  CAB_PROF_T(t_kernel)
  extern __shared__ __align__(1024) uint8_t smem[];
  const int n_stages = args.n_stages, stage_bytes = args.stage_bytes;
  const int n_warps = (int) (blockDim.x >> 5), n_pairs = n_warps >> 1;
  uint8_t *ring = smem;
  double2 *abuf_all = reinterpret_cast<double2 *>(smem + fwd_smem_ring(n_stages, stage_bytes));
  ChunkDesc *s_chunks = reinterpret_cast<ChunkDesc *>(smem + fwd_smem_ring(n_stages, stage_bytes) + fwd_smem_abuf(n_pairs, args.a_units));
  double *exp_tbl = reinterpret_cast<double *>(s_chunks + kSmemChunks);  // 2^(j/256) for the table-driven exp
  double *dot_scratch = exp_tbl + kExpTblSize;                           // [pair][8]: half 1's partial sums of the fused output layer
  uint64_t *full = reinterpret_cast<uint64_t *>(dot_scratch + kDotScratchBytes / sizeof(double));
  uint64_t *tbl_bar = full + n_stages;                                   // completes when the tables have landed
  uint32_t *done = reinterpret_cast<uint32_t *>(tbl_bar + 1);            // arrivals per ring slot, monotonically increasing

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = warp >> 1, half = pair_half(warp);
  const int m = lane >> 2, q = lane & 3;
  // chunks this CTA will consume, in order: (groups it owns) x n_chunks
  const int my_groups = args.n_groups > blockIdx.x ? (int) ((args.n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  const int total_chunks = my_groups * args.n_chunks;   // (a CTA never sees 2^31 chunks: 2^20 groups of <= 2^11 chunks)
  const bool table_in_smem = args.n_chunks <= kSmemChunks;
  const ChunkDesc *chunks = table_in_smem ? s_chunks : args.chunks;

  // A CTA of a 4096-board launch lives for ONE pass, so its start-up is part of every pass: the first group's input
  // codes are requested before anything else (their DRAM latency runs under the barrier set-up), the chunk and exp
  // tables arrive by one bulk copy instead of per-thread loads, and the first ring fills need no global read.
  LaneCodes codes_first;
  {
    const long board = ((long) blockIdx.x * n_pairs + pair) * kBoardsPerWarp + m;
    codes_first = load_codes(args.codes, board, my_groups > 0 && board < args.n_boards, half, q);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&full[s], 1);
      done[s] = 0;
    }
    mbar_init(tbl_bar, 1);
    mbar_fence_init();
    for (int s = 0; s < n_stages && s < total_chunks; ++s) {  // initial fill of the ring
      mbar_arrive_expect_tx(&full[s], args.fill_bytes[s]);
      bulk_g2s(ring + (size_t) s * stage_bytes, args.pack + args.fill_off[s], args.fill_bytes[s], &full[s]);
    }
    const uint32_t tb = table_in_smem ? (uint32_t) kBlobBytes : (uint32_t) (kExpTblSize * sizeof(double));
    mbar_arrive_expect_tx(tbl_bar, tb);
    if (table_in_smem) bulk_g2s(s_chunks, args.blob, tb, tbl_bar);
    else bulk_g2s(exp_tbl, args.blob + kBlobChunkBytes, tb, tbl_bar);
  }
  __syncthreads();

  const long long t_start = args.team_lag > 0 ? clock64() : 0;
  double2 *abuf = abuf_all + (size_t) pair * args.a_units * 32;
  double acc[kMaxNBH][2];
  uint32_t rs = 0, rk = 0;   // ring slot of the current chunk, times the ring wrapped
  int n = 0;                 // sequence number of the current chunk within this CTA
  const uint32_t abuf_addr = smem_u32(abuf) + (uint32_t) lane * 16u;
  const uint32_t ring_addr = smem_u32(ring) + (uint32_t) lane * 16u;

  for (long g = blockIdx.x; g < args.n_groups; g += gridDim.x) {
    const long board0 = (g * n_pairs + pair) * kBoardsPerWarp;
    const long board = board0 + m;
    const bool valid = board < args.n_boards;
    const bool first_group = g == (long) blockIdx.x;
    // K2: this lane's 16 bytes inside unit row 0 of the saved-activation tensor (unit layout [unit][board][8])
    double *save_lane = (SAVE && args.save != nullptr && valid) ? args.save + board0 * 8 + lane * 2 : nullptr;

    CAB_PROF_T(t_load)
    {
      // Network::loadBoard (network.cpp:119-123): input = k * 0.4, bit-identical to Piece::code().
      // The pair splits the 8 input units; the previous group's last reads of abuf ended at its final pair barrier.
      const LaneCodes lc = first_group ? codes_first : load_codes(args.codes, board, valid, half, q);
#pragma unroll
      for (int tt = 0; tt < 4; ++tt) {
        const int t = 2 * tt + half;
        const uint32_t pk = lc.v[tt >> 1] >> (16 * (tt & 1));
        const int c0 = (int) (int8_t) (pk & 0xffu), c1 = (int) (int8_t) ((pk >> 8) & 0xffu);
        const double2 x = make_double2((double) c0 * 0.4, (double) c1 * 0.4);
        abuf[t * 32 + lane] = x;
        if (SAVE && save_lane != nullptr) *reinterpret_cast<double2 *>(save_lane + (size_t) t * args.save_stride) = x;
      }
      pair_barrier(pair);
      if (first_group) mbar_wait(tbl_bar, 0);   // chunk and exp tables present (requested before the input load)
    }
    CAB_PROF_ACC(0, t_load)
    // Optional staggered teams (CAB_FWD_LAG; measured neutral for one-pass CTAs: the delayed team's gain equals the delay).
    if (args.team_lag > 0 && first_group) {
      const int team = warp >> args.team_shift;
      if (team > 0) {
        const long long until = t_start + (long long) team * args.team_lag;
        while (clock64() < until) {}
      }
    }

    for (int c = 0; c < args.n_chunks; ++c, ++n) {
      CAB_PROF_T(t_desc)
      const ChunkDesc d = chunks[c];
      const bool dot = d.flags & kDot;
      const bool first = (d.flags & kFirst) && !dot, last = (d.flags & kLast) && !dot;
      int j0, nbh;
      // a fused output layer works on the PREVIOUS layer's tiles (its input units), split like that layer was
      split_tiles(dot ? d.tcount : d.nb, dot ? d.layer - 1 : d.layer, half, j0, nbh);
      if (first) {
#pragma unroll
        for (int j = 0; j < kMaxNBH; ++j) acc[j][0] = acc[j][1] = 0.0;
      }
      CAB_PROF_ACC(1, t_desc)
      CAB_PROF_T(t_wait)
      mbar_wait(&full[rs], rk & 1);
      CAB_PROF_ACC(2, t_wait)
      CAB_PROF_T(t_mma)
      double dot_sum = 0.0;
      if (!dot) {
        dispatch_chunk(nbh, acc, abuf_addr + (uint32_t) (d.a_in + d.t0) * kUnitBytes,
                       ring_addr + rs * (uint32_t) stage_bytes + (uint32_t) j0 * kUnitBytes, (uint32_t) d.nb * kUnitBytes, d.tcount);
      } else {
        // One-output layer (network.cpp:99-114 with N = 1) on the activations this warp still holds in registers:
        // lane (m, q) owns features 8 (j0 + j) + 2q + {0, 1} of board m; the chunk's unit t keeps W[8t + 2q' + h][0]
        // in lane q' (B fragment of output column 0), so all lanes of one q read the same 16 bytes.
        dot_sum = dispatch_dot(nbh, acc, smem_u32(ring) + rs * (uint32_t) stage_bytes + (uint32_t) j0 * kUnitBytes + (uint32_t) q * 16u);
        dot_sum += __shfl_xor_sync(0xffffffffu, dot_sum, 1);
        dot_sum += __shfl_xor_sync(0xffffffffu, dot_sum, 2);
      }
      __syncwarp();
      CAB_PROF_ACC(3, t_mma)

      CAB_PROF_T(t_refill)
      // Release the slot. In-order issue means every load of this warp from the slot has returned before the
      // atomic issues; the last warp to arrive therefore re-arms the slot with the chunk that is n_stages ahead.
      if (lane == 0) {
        const uint32_t old = atomicAdd(&done[rs], 1u);
        if ((old + 1) % (uint32_t) n_warps == 0) {
          const int refill = n + n_stages;
          if (refill < total_chunks) {
            const ChunkDesc *nd = &chunks[refill % args.n_chunks];
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive_expect_tx(&full[rs], nd->bytes);
            bulk_g2s(ring + (size_t) rs * stage_bytes, args.pack + nd->src_off, nd->bytes, &full[rs]);
          }
        }
      }
      if (++rs == (uint32_t) n_stages) { rs = 0; ++rk; }
      CAB_PROF_ACC(4, t_refill)

      CAB_PROF_T(t_act)
      if (dot) {
        // the pair's two partial sums meet in shared memory; half 0 finishes the board: theta, y = f(theta), outputs
        double *scratch = dot_scratch + pair * kBoardsPerWarp;
        if (half == 1 && q == 0) scratch[m] = dot_sum;
        pair_barrier(pair);
        if (half == 0) {
          const double theta = dot_sum + scratch[m];
          const double y = act_f_tbl(theta, exp_tbl);
          const size_t goff = (size_t) board0 * 8 + lane * 2 + (size_t) d.save_off * args.save_stride;
          if (THETA && args.save_theta != nullptr && valid)
            *reinterpret_cast<double2 *>(args.save_theta + goff) = make_double2(q == 0 ? theta : 0.0, 0.0);
          if (SAVE && args.save != nullptr && valid) *reinterpret_cast<double2 *>(args.save + goff) = make_double2(q == 0 ? y : 0.0, 0.0);
          if (args.out != nullptr && valid && q == 0) args.out[board] = y;
        }
      } else if (last) {
        // Layer boundary. (1) a = f(theta) on this warp's own tiles, IN REGISTERS and BEFORE the barrier: the wait for
        // the partner (who may still be reading the layer input that the outputs overwrite in place) is spent on the
        // exp / reciprocal math, and there is no shared-memory round trip. (2) barrier: input fully read.
        // (3) store the activations in place. (4) barrier: outputs visible to the partner.
        // In front of a fused output layer (kPreDot) the activations simply stay in registers: no barrier, no store;
        // K2 / theta export write them to global memory straight from the registers.
        const bool keep = d.flags & kPreDot;
        double2 *a_out = abuf + (size_t) (d.a_out + j0) * 32;
        if (THETA && args.save_theta != nullptr && valid)
          dispatch_store_global(nbh, acc, args.save_theta + board0 * 8 + lane * 2 + (size_t) (d.save_off + d.tile0 + j0) * args.save_stride,
                                args.save_stride);
        act_all(nbh, acc, exp_tbl);
        if (SAVE && save_lane != nullptr)                // K2: keep every layer's activations for the backward pass
          dispatch_store_global(nbh, acc, save_lane + (size_t) (d.save_off + d.tile0 + j0) * args.save_stride, args.save_stride);
        if (!keep) {
          CAB_PROF_T(t_bar1)
          pair_barrier(pair);
          CAB_PROF_ACC(7, t_bar1)
          dispatch_store(nbh, acc, a_out, lane);
          const bool owns_tile0 = j0 == 0 && nbh > 0;
          if ((d.flags & kFinal) && owns_tile0 && args.out != nullptr && valid && q == 0) args.out[board] = a_out[lane].x;
          CAB_PROF_T(t_bar2)
          pair_barrier(pair);
          CAB_PROF_ACC(7, t_bar2)
        }
      }
      CAB_PROF_ACC(5, t_act)
    }
  }
  if (PROF && args.prof != nullptr && lane == 0) {
    prof_acc[6] = clock64() - t_kernel;
    long long *dst = args.prof + ((long) blockIdx.x * n_warps + warp) * 8;
    for (int i = 0; i < 8; ++i) dst[i] = prof_acc[i];
  }
  // every issued bulk copy has been waited on by every warp (total_chunks bounds the refills): nothing in flight at exit
}

}  // namespace cab
