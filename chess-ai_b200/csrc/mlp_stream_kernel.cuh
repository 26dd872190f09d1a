// mlp_stream_kernel.cuh — the fused whole-MLP forward kernel (K1 / K2).
//
// Replaces Network::predictPosition / propagate_for_training (/root/reference/src/mcts_network/network.cpp:92-153):
//   per layer, per output j:  sum_i a[i] * W[i][j]  ->  f(sum)
// B200 design:
//   * one warp owns 8 boards through ALL layers; activations never leave its private shared-memory units
//   * the matmuls run on the fp64 tensor pipe: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4), accumulators in registers
//   * weights arrive as a pre-packed B-fragment stream, moved L2 -> shared by 1-D TMA bulk copies
//     (cp.async.bulk, SASS UBLKCP) issued by a producer warp into a 3-slot mbarrier ring shared by the CTA
//   * the DMMA C-fragment of tile j IS the A-fragment pair of the next layer's k-steps (2j, 2j+1) under the feature
//     permutation k = 8t + 2q + h that the packed stream bakes in, so there is no transpose, no __syncthreads and
//     no inter-warp traffic between layers
//   * persistent grid: CTAs loop over 32-board groups; the weight stream is re-read from L2 per group
#pragma once
#include "cab_internal.cuh"

namespace cab {

template <int NB>
__device__ __forceinline__ void run_chunk(double (&acc)[kMaxNB][2], const double2 *__restrict__ a_in,
                                          const double2 *__restrict__ w, int tcount, int lane) {
  for (int t = 0; t < tcount; ++t) {
    const double2 a = a_in[t * 32 + lane];
    const double2 *wt = w + (size_t) t * NB * 32 + lane;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const double2 b = wt[j * 32];
      dmma884(acc[j], a.x, b.x);
      dmma884(acc[j], a.y, b.y);
    }
  }
}

template <int NB>
__device__ __forceinline__ void store_acc(const double (&acc)[kMaxNB][2], double2 *__restrict__ a_out, int lane) {
#pragma unroll
  for (int j = 0; j < NB; ++j) a_out[j * 32 + lane] = make_double2(acc[j][0], acc[j][1]);
}

#define CAB_NB_CASE(N)                        \
  case N:                                     \
    run_chunk<N>(acc, a_in, w, tcount, lane); \
    if (last) store_acc<N>(acc, a_out, lane); \
    break;

__device__ __forceinline__ void dispatch_chunk(int nb, double (&acc)[kMaxNB][2], const double2 *a_in, const double2 *w,
                                               int tcount, int lane, bool last, double2 *a_out) {
  switch (nb) {
    CAB_NB_CASE(1) CAB_NB_CASE(2) CAB_NB_CASE(3) CAB_NB_CASE(4) CAB_NB_CASE(5) CAB_NB_CASE(6) CAB_NB_CASE(7)
    CAB_NB_CASE(8) CAB_NB_CASE(9) CAB_NB_CASE(10) CAB_NB_CASE(11) CAB_NB_CASE(12) CAB_NB_CASE(13) CAB_NB_CASE(14)
    CAB_NB_CASE(15) CAB_NB_CASE(16) CAB_NB_CASE(17) CAB_NB_CASE(18) CAB_NB_CASE(19) CAB_NB_CASE(20) CAB_NB_CASE(21)
    CAB_NB_CASE(22) CAB_NB_CASE(23) CAB_NB_CASE(24) CAB_NB_CASE(25) CAB_NB_CASE(26) CAB_NB_CASE(27) CAB_NB_CASE(28)
    CAB_NB_CASE(29) CAB_NB_CASE(30) CAB_NB_CASE(31) CAB_NB_CASE(32)
    default: break;
  }
}
#undef CAB_NB_CASE

struct ForwardArgs {
  const uint8_t *pack;      // packed fragment stream of W
  const ChunkDesc *chunks;
  int n_chunks;
  int a_units;              // per-warp activation buffer, in units
  long n_boards;
  long n_groups;            // ceil(n_boards / (8 * consumer warps))
  const int8_t *codes;      // [n_boards][64]
  double *out;              // [n_boards] or nullptr
  double *save;             // saved activations in unit layout, or nullptr (K2: propagate_for_training)
  long save_stride;         // doubles per unit row = boards_pad * 8
};

__global__ void __launch_bounds__(kThreads, 2) mlp_forward_kernel(const ForwardArgs args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *ring = smem;
  double2 *abuf_all = reinterpret_cast<double2 *>(smem + kStages * kStageBytes);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + kStages * kStageBytes +
                                                (size_t) ((blockDim.x >> 5) - 1) * args.a_units * kUnitBytes);
  uint64_t *empty = full + kStages;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_cons = (int) (blockDim.x >> 5) - 1;  // consumer warps in this launch (<= kConsumerWarps)
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], n_cons);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == n_cons) {
    // ---------------- producer: stream the packed weights into the ring, once per board group ----------------
    if (lane == 0) {
      uint32_t cc = 0;
      for (long g = blockIdx.x; g < args.n_groups; g += gridDim.x) {
        for (int c = 0; c < args.n_chunks; ++c, ++cc) {
          const uint32_t src_off = args.chunks[c].src_off, bytes = args.chunks[c].bytes;
          const uint32_t s = cc % kStages, k = cc / kStages;
          if (k > 0) mbar_wait(&empty[s], (k - 1) & 1);
          mbar_arrive_expect_tx(&full[s], bytes);
          bulk_g2s(ring + (size_t) s * kStageBytes, args.pack + src_off, bytes, &full[s]);
        }
      }
    }
    return;
  }

  // ---------------- consumers ------------------------------------------------------------------------------
  double2 *abuf = abuf_all + (size_t) warp * args.a_units * 32;
  const int m = lane >> 2, q = lane & 3;
  double acc[kMaxNB][2];
  uint32_t cc = 0;

  for (long g = blockIdx.x; g < args.n_groups; g += gridDim.x) {
    const long board0 = (g * n_cons + warp) * kBoardsPerWarp;
    const long board = board0 + m;
    const bool valid = board < args.n_boards;

    {
      // Network::loadBoard (network.cpp:119-123): input = k * 0.4, bit-identical to Piece::code()
      const int8_t *row = args.codes + (valid ? board : 0) * 64;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int c0 = valid ? row[8 * t + 2 * q] : 0, c1 = valid ? row[8 * t + 2 * q + 1] : 0;
        const double2 x = make_double2((double) c0 * 0.4, (double) c1 * 0.4);
        abuf[t * 32 + lane] = x;
        if (args.save != nullptr && valid)
          *reinterpret_cast<double2 *>(args.save + (size_t) t * args.save_stride + board0 * 8 + lane * 2) = x;
      }
    }

    for (int c = 0; c < args.n_chunks; ++c, ++cc) {
      const ChunkDesc d = args.chunks[c];
      const bool first = d.flags & kFirst, last = d.flags & kLast;
      if (first) {
#pragma unroll
        for (int j = 0; j < kMaxNB; ++j) acc[j][0] = acc[j][1] = 0.0;
      }
      const uint32_t s = cc % kStages, k = cc / kStages;
      mbar_wait(&full[s], k & 1);
      const double2 *w = reinterpret_cast<const double2 *>(ring + (size_t) s * kStageBytes);
      double2 *a_out = abuf + (size_t) d.a_out * 32;
      dispatch_chunk(d.nb, acc, abuf + (size_t) (d.a_in + d.t0) * 32, w, d.tcount, lane, last, a_out);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);

      if (last) {
        // a = f(theta) on this n-block, in place on the units this lane wrote itself (no cross-lane hazard)
#pragma unroll 2
        for (int j = 0; j < d.nb; ++j) {
          double2 v = a_out[j * 32 + lane];
          v.x = act_f(v.x);
          v.y = act_f(v.y);
          a_out[j * 32 + lane] = v;
          if (args.save != nullptr && valid)
            *reinterpret_cast<double2 *>(args.save + (size_t) (d.save_off + d.tile0 + j) * args.save_stride +
                                         board0 * 8 + lane * 2) = v;
        }
        if ((d.flags & kFinal) && args.out != nullptr && valid && q == 0) args.out[board] = a_out[lane].x;
      }
    }
  }
}

}  // namespace cab
