// tc_gemm_bf16.cuh — bf16 tensor-core path for WIDE layers (K8; BASELINE.json configs[4]: widened MLP 4 x 2048).
//
//   Y[M][N] (bf16) = f( X[M][K] (bf16, K contiguous) * Wt[N][K]^T (bf16, K contiguous) )      fp32 accumulate
//
// Same math as one layer of Network::predictPosition (/root/reference/src/mcts_network/network.cpp:99-114) at the
// stated bf16 tolerance; used only where the batched layer really is a dense contraction (width a multiple of 256).
// The default 64-144-225-100-1 network never takes this path.
//
// Blackwell structure (one CTA per SM, persistent over output tiles):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d (SASS UTMALDG) of a 128x64 X tile and a 256x64 Wt tile per
//               k-block into a 4-stage shared-memory ring, 128-byte swizzle, completion on "full" mbarriers
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (SASS UTCHMMA), M=128 N=256
//               K=16, four per k-block, accumulating into TMEM; tcgen05.commit releases the smem stage ("empty")
//               and, after the last k-block, publishes the accumulator ("tmem_full")
//   warp 2      allocates / frees the 512 TMEM columns (two 256-column accumulators: the epilogue of tile i overlaps
//               the main loop of tile i+1)
//   warps 4-7   epilogue: tcgen05.ld (SASS LDTM) 32 lanes x 32 columns at a time, f(x) = 4*tanh(x/4) in fp32,
//               convert to bf16, 64-byte row segments to global; then arrive on "tmem_empty"
#pragma once
#include <cuda_bf16.h>

#include "cab_internal.cuh"

namespace cab {
namespace tc {

constexpr int BM = 128, BN = 256, BK = 64;      // CTA tile; BK * 2 B = 128 B = one swizzle atom row
constexpr int UMMA_K = 16;
constexpr int kStagesTc = 4;
constexpr int kAccumStages = 2;
constexpr int kTmemCols = 512;
constexpr int kThreadsTc = 256;
constexpr uint32_t kBytesA = BM * BK * 2, kBytesB = BN * BK * 2;
constexpr size_t kSmemTc = (size_t) kStagesTc * (kBytesA + kBytesB) + 1024 /*align slack*/ + 256 /*barriers*/;

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (bit 4), a/b format BF16 (bits 7, 10),
// K-major A and B (bits 15, 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (BN >> 3) << 17) | ((uint32_t) (BM >> 4) << 24);

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor) for a K-major tile with 128-byte swizzle:
// start address >> 4, LBO unused (one swizzle atom along K), SBO = 8 rows * 128 B = 1024 B, version 1, layout 2
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t) ((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t) (1024 >> 4) << 32;
  d |= (uint64_t) 1 << 46;
  d |= (uint64_t) 2 << 61;
  return d;
}

__device__ __forceinline__ void tma_load_2d(void *dst, const void *tmap, int c0, int c1, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// f(x) = 4 * tanh(x / 4) (network.h:83-94 in closed form), fp32 with the hardware tanh: adequate for a bf16 result
__device__ __forceinline__ float act_f32(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * 0.25f));
  return 4.0f * t;
}

// Epilogue flavours of the one GEMM kernel  D[M][N] = A[M][K] * Bt[N][K]^T  (both operands contraction-contiguous):
//   kEpiAct    forward layer:   Y = f(D)                        bf16 row-major (+ optional transposed copy Yt[N][ldt])
//   kEpiDeriv  backward dX:     Y = D * f'(a), a = aux[M][N]    bf16 row-major (+ transposed copy): psi of the layer
//                               below (network.cpp:250-257: omega accumulated through W, times df of the saved theta;
//                               f'(theta) = 1 - (a/4)^2 in closed form from the saved activation)
//   kEpiGrad   backward dW:     G[M][N] (+)= D                  fp64, row-major = the reference's [in][out] weight order
//                               (network.cpp:258: dW = psi * a summed over the batch); split-K work items add with red.f64
enum { kEpiAct = 0, kEpiDeriv = 1, kEpiGrad = 2 };

struct TcGemmArgs {
  int M, N, K;
  __nv_bfloat16 *Y;          // [M][N]                      (kEpiAct / kEpiDeriv)
  __nv_bfloat16 *Yt;         // [N][ldt] or nullptr         (kEpiAct / kEpiDeriv)
  long ldt;
  const __nv_bfloat16 *aux;  // [M][N] saved activation     (kEpiDeriv)
  double *G;                 // [M][N]                      (kEpiGrad)
  int split_k;               // work items per output tile  (kEpiGrad; 1 elsewhere)
  int apply_act;             // kEpiAct: 0 = plain product
};

// One 32-column chunk of one accumulator row (r[] = the thread's 32 fp32 values from TMEM) through the epilogue EPI.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const TcGemmArgs &args, const uint32_t (&r)[32], int row, size_t row_off, int nb, int c,
                                               int n_splits, int lane) {
  if (EPI == kEpiGrad) {
    if (row < args.M) {
      double *g = args.G + row_off + c;
      if (n_splits == 1) {
#pragma unroll
        for (int i = 0; i < 32; i += 2)
          *reinterpret_cast<double2 *>(g + i) = make_double2((double) __uint_as_float(r[i]), (double) __uint_as_float(r[i + 1]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          asm volatile("red.global.add.f64 [%0], %1;" ::"l"(g + i), "d"((double) __uint_as_float(r[i])) : "memory");
      }
    }
  } else {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
    if (EPI == kEpiAct) {
      if (args.apply_act) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = act_f32(v[i]);
      }
    } else {
      uint4 a4[4] = {};
      if (row < args.M) {
        const uint4 *ap = reinterpret_cast<const uint4 *>(args.aux + row_off + c);
#pragma unroll
        for (int i = 0; i < 4; ++i) a4[i] = __ldg(ap + i);
      }
      const __nv_bfloat162 *ab = reinterpret_cast<const __nv_bfloat162 *>(a4);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 a = __bfloat1622float2(ab[i]);
        v[2 * i] *= 1.0f - 0.0625f * a.x * a.x;
        v[2 * i + 1] *= 1.0f - 0.0625f * a.y * a.y;
      }
    }
    if (row >= args.M) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0.0f;  // pad rows of the transposed copy stay zero
    }
    if (row < args.M) {
      uint4 *dst = reinterpret_cast<uint4 *>(args.Y + row_off + c);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __nv_bfloat162 b = __floats2bfloat162_rn(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]);
          pk[j] = *reinterpret_cast<uint32_t *>(&b);
        }
        dst[i] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
    if (args.Yt != nullptr && row < args.ldt) {   // warp-uniform: a warp holds 32 consecutive rows and ldt is a multiple of 256
      // transposed copy: lanes are consecutive rows; lane pairs swap so that each lane stores two rows of one
      // column as one 4-byte word -> every column receives 64 contiguous bytes from the warp
      __nv_bfloat16 *tcol = args.Yt + (size_t) (nb * BN + c) * args.ldt + (row & ~1);
      const bool odd = lane & 1;
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float send = odd ? v[i] : v[i + 1];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
        const __nv_bfloat162 b = odd ? __floats2bfloat162_rn(recv, v[i + 1]) : __floats2bfloat162_rn(v[i], recv);
        *reinterpret_cast<__nv_bfloat162 *>(tcol + (size_t) (i + (odd ? 1 : 0)) * args.ldt) = b;
      }
    }
  }
}

template <int EPI>
__global__ void __launch_bounds__(kThreadsTc, 1)
tc_gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const TcGemmArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);  // swizzle-128B atoms
  uint8_t *smem_a = smem;
  uint8_t *smem_b = smem + (size_t) kStagesTc * kBytesA;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t) kStagesTc * (kBytesA + kBytesB));
  uint64_t *empty = full + kStagesTc;
  uint64_t *tmem_full = empty + kStagesTc;
  uint64_t *tmem_empty = tmem_full + kAccumStages;
  uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(tmem_empty + kAccumStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blocks_n = args.N / BN, n_blocks_m = (args.M + BM - 1) / BM;
  const int n_kblocks = (args.K + BK - 1) / BK;            // a ragged tail is zero-filled by TMA
  const int kb_per_item = (n_kblocks + args.split_k - 1) / args.split_k;
  const int n_splits = (n_kblocks + kb_per_item - 1) / kb_per_item;
  const int n_items = n_blocks_m * n_blocks_n * n_splits;  // item = (tile, k-split), splits of a tile adjacent

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStagesTc; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < kAccumStages; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
    mbar_fence_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ------------------------------------------- TMA producer -------------------------------------------
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int tile = item / n_splits, split = item % n_splits;
        const int mb = tile / n_blocks_n, nb = tile % n_blocks_n;
        const int kb0 = split * kb_per_item, kb1 = min(kb0 + kb_per_item, n_kblocks);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], kBytesA + kBytesB);
          tma_load_2d(smem_a + (size_t) s * kBytesA, &tmap_x, kb * BK, mb * BM, &full[s]);
          tma_load_2d(smem_b + (size_t) s * kBytesB, &tmap_w, kb * BK, nb * BN, &full[s]);
          if (++s == kStagesTc) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------- MMA issuer ---------------------------------------------
    uint32_t s = 0, ph = 0, as = 0, aph = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int split = item % n_splits;
      const int kb0 = split * kb_per_item, kb1 = min(kb0 + kb_per_item, n_kblocks);
      mbar_wait(&tmem_empty[as], aph ^ 1);          // the epilogue drained this accumulator
      tmem_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t) as * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[s], ph);                    // TMA bytes landed
        tmem_fence_after();
        if (lane == 0) {
          const uint64_t a_desc = make_smem_desc(smem_u32(smem_a + (size_t) s * kBytesA));
          const uint64_t b_desc = make_smem_desc(smem_u32(smem_b + (size_t) s * kBytesB));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)     // advance 32 B inside the swizzle atom: +2 in the >>4 address field
            umma_f16(tmem_d, a_desc + (uint64_t) (k * 2), b_desc + (uint64_t) (k * 2), kIdesc, (kb != kb0 || k != 0) ? 1u : 0u);
          umma_commit(&empty[s]);                   // frees the smem stage when these MMAs retire
          if (kb == kb1 - 1) umma_commit(&tmem_full[as]);
        }
        __syncwarp();
        if (++s == kStagesTc) { s = 0; ph ^= 1; }
      }
      if (++as == kAccumStages) { as = 0; aph ^= 1; }
    }
  } else if (warp >= 4) {
    // ------------------------------------------- epilogue -----------------------------------------------
    const int q = warp & 3;                          // TMEM lane quarter this warp may read
    uint32_t as = 0, aph = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int tile = item / n_splits;
      const int mb = tile / n_blocks_n, nb = tile % n_blocks_n;
      mbar_wait(&tmem_full[as], aph);
      tmem_fence_after();
      const int row = mb * BM + q * 32 + lane;
      const size_t row_off = (size_t) row * args.N + (size_t) nb * BN;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t) (q * 32) << 16) + (uint32_t) (as * BN + c), r);
        epilogue_chunk<EPI>(args, r, row, row_off, nb, c, n_splits, lane);
      }
      tmem_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      if (++as == kAccumStages) { as = 0; aph ^= 1; }
    }
  }

  tmem_fence_before();
  __syncthreads();
  if (warp == 2) {
    tmem_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-PAIR variant (cta_group::2): two CTAs of one cluster (the two SMs of a TPC) compute a 256 x 256 output tile.
// Each CTA stages ITS 128 rows of A and ITS 128 of the 256 B rows (32 KB per k-block instead of 48 KB: the single-CTA
// kernel needs ~64 B/clk/SM from L2 at full tensor rate, more than L2 delivers to 148 SMs at once), the leader CTA
// issues one tcgen05.mma.cta_group::2 with M = 256 that reads both CTAs' shared memory and writes each CTA's half of
// the accumulator into that CTA's own TMEM; every CTA runs its own epilogue on its 128 rows.
//   full[s]        waited by the leader's MMA warp only; BOTH CTAs' TMA loads complete_tx on the LEADER's barrier
//                  (barrier address with the peer bit cleared), the leader expects the bytes of both
//   empty[s]       per CTA; the leader's tcgen05.commit multicasts the arrival to both CTAs
//   tmem_full[a]   per CTA, same multicast commit
//   tmem_empty[a]  the leader's; 8 arrivals = the 4 epilogue warps of each CTA (the peer's arrive through mapa)
// ---------------------------------------------------------------------------------------------------------------
constexpr int kStagesTc2 = 6;
constexpr uint32_t kBytesA2 = BM * BK * 2, kBytesB2 = (BN / 2) * BK * 2;
constexpr size_t kSmemTc2 = (size_t) kStagesTc2 * (kBytesA2 + kBytesB2) + 1024 + 256;
constexpr uint32_t kIdesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (BN >> 3) << 17) | ((uint32_t) ((2 * BM) >> 4) << 24);
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address inside a CTA pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const void *tmap, int c0, int c1, uint64_t *leader_bar_local) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(leader_bar_local) & kPeerBitMask), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {   // arrives on the same barrier offset in both CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t) 3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar_local) {   // arrive on CTA 0's copy of this barrier
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar_local))
      : "memory");
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsTc, 1)
tc_gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const TcGemmArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
  uint8_t *smem_a = smem;
  uint8_t *smem_b = smem + (size_t) kStagesTc2 * kBytesA2;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t) kStagesTc2 * (kBytesA2 + kBytesB2));
  uint64_t *empty = full + kStagesTc2;
  uint64_t *tmem_full = empty + kStagesTc2;
  uint64_t *tmem_empty = tmem_full + kAccumStages;
  uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(tmem_empty + kAccumStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_blocks_n = args.N / BN, n_blocks_m = (args.M + 2 * BM - 1) / (2 * BM);   // 256-row blocks
  const int n_kblocks = (args.K + BK - 1) / BK;
  const int kb_per_item = (n_kblocks + args.split_k - 1) / args.split_k;
  const int n_splits = (n_kblocks + kb_per_item - 1) / kb_per_item;
  const int n_items = n_blocks_m * n_blocks_n * n_splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStagesTc2; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < kAccumStages; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 8); }
    mbar_fence_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tmem_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs' barriers exist before anybody signals across the pair
  tmem_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ------------------------------------------- TMA producer (both CTAs) -------------------------------
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (int item = pair_id; item < n_items; item += n_pairs) {
        const int tile = item / n_splits, split = item % n_splits;
        const int mb = tile / n_blocks_n, nb = tile % n_blocks_n;
        const int kb0 = split * kb_per_item, kb1 = min(kb0 + kb_per_item, n_kblocks);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * (kBytesA2 + kBytesB2));
          tma_load_2d_pair(smem_a + (size_t) s * kBytesA2, &tmap_x, kb * BK, mb * 2 * BM + (int) rank * BM, &full[s]);
          tma_load_2d_pair(smem_b + (size_t) s * kBytesB2, &tmap_w, kb * BK, nb * BN + (int) rank * (BN / 2), &full[s]);
          if (++s == kStagesTc2) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------------------- MMA issuer (leader CTA only) ---------------------------
    uint32_t s = 0, ph = 0, as = 0, aph = 0;
    for (int item = pair_id; item < n_items; item += n_pairs) {
      const int split = item % n_splits;
      const int kb0 = split * kb_per_item, kb1 = min(kb0 + kb_per_item, n_kblocks);
      mbar_wait(&tmem_empty[as], aph ^ 1);          // both CTAs' epilogues drained this accumulator
      tmem_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t) as * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[s], ph);                    // both CTAs' bytes landed
        tmem_fence_after();
        if (lane == 0) {
          const uint64_t a_desc = make_smem_desc(smem_u32(smem_a + (size_t) s * kBytesA2));
          const uint64_t b_desc = make_smem_desc(smem_u32(smem_b + (size_t) s * kBytesB2));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_f16_pair(tmem_d, a_desc + (uint64_t) (k * 2), b_desc + (uint64_t) (k * 2), kIdesc2, (kb != kb0 || k != 0) ? 1u : 0u);
          umma_commit_pair(&empty[s]);              // frees the stage in both CTAs
          if (kb == kb1 - 1) umma_commit_pair(&tmem_full[as]);
        }
        __syncwarp();
        if (++s == kStagesTc2) { s = 0; ph ^= 1; }
      }
      if (++as == kAccumStages) { as = 0; aph ^= 1; }
    }
  } else if (warp >= 4) {
    // ------------------------------------------- epilogue (both CTAs, own 128 rows) ---------------------
    const int q = warp & 3;
    uint32_t as = 0, aph = 0;
    for (int item = pair_id; item < n_items; item += n_pairs) {
      const int tile = item / n_splits;
      const int mb = tile / n_blocks_n, nb = tile % n_blocks_n;
      mbar_wait(&tmem_full[as], aph);
      tmem_fence_after();
      const int row = mb * 2 * BM + (int) rank * BM + q * 32 + lane;
      const size_t row_off = (size_t) row * args.N + (size_t) nb * BN;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t) (q * 32) << 16) + (uint32_t) (as * BN + c), r);
        epilogue_chunk<EPI>(args, r, row, row_off, nb, c, n_splits, lane);
      }
      tmem_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[as]);
      if (++as == kAccumStages) { as = 0; aph ^= 1; }
    }
  }

  tmem_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer's shared memory and barriers stay alive until the leader is done with them
  if (warp == 2) {
    tmem_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
}

// int8 codes -> bf16 inputs k * 0.4  (Network::loadBoard at bf16 precision)
__global__ void __launch_bounds__(256) codes_to_bf16_kernel(const int8_t *__restrict__ codes, long n, __nv_bfloat16 *__restrict__ x) {
  for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x)
    x[i] = __float2bfloat16_rn((float) codes[i] * 0.4f);
}

// last layer (N = 1): out[m] = f( sum_k Y[m][k] * w[k] ), one warp per board, fp32 accumulate
__global__ void __launch_bounds__(256) gemv_out_kernel(const __nv_bfloat16 *__restrict__ y, const float *__restrict__ w, long M, int K,
                                                        double *__restrict__ out) {
  const long row = (blockIdx.x * (long) blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const __nv_bfloat162 *yr = reinterpret_cast<const __nv_bfloat162 *>(y + row * K);
  float acc = 0.f;
  for (int k = lane; k < K / 2; k += 32) {
    const float2 v = __bfloat1622float2(yr[k]);
    acc = fmaf(v.x, w[2 * k], acc);
    acc = fmaf(v.y, w[2 * k + 1], acc);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[row] = (double) act_f32(acc);
}

// fp64 weights W[K][N] -> bf16 Wt[N][K]
__global__ void __launch_bounds__(256) pack_wt_bf16_kernel(const double *__restrict__ w, int K, int N, __nv_bfloat16 *__restrict__ wt) {
  const long total = (long) K * N;
  for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < total; i += (long) gridDim.x * blockDim.x) {
    const int n = (int) (i / K), k = (int) (i % K);
    wt[i] = __float2bfloat16_rn((float) w[(long) k * N + n]);
  }
}
__global__ void __launch_bounds__(256) pack_last_f32_kernel(const double *__restrict__ w, int K, float *__restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K; i += gridDim.x * blockDim.x) out[i] = (float) w[i];
}


// ---- training-side helpers of the bf16 path (all HBM-trivial next to the GEMMs) ---------------------------------
// transposed copy of the inputs: xt[j][b] = k * 0.4 (row stride ldt, pad columns zeroed by the caller's memset)
__global__ void __launch_bounds__(256) codes_to_bf16_t_kernel(const int8_t *__restrict__ codes, long n_boards, long ldt,
                                                               __nv_bfloat16 *__restrict__ xt) {
  const long total = n_boards * 64;
  for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < total; i += (long) gridDim.x * blockDim.x) {
    const long j = i / n_boards, b = i % n_boards;
    xt[j * ldt + b] = __float2bfloat16_rn((float) codes[b * 64 + j] * 0.4f);
  }
}

// W[K][N] fp64 -> bf16, same layout: the B operand of the dX GEMM (contraction over N)
__global__ void __launch_bounds__(256) pack_w_bf16_kernel(const double *__restrict__ w, long total, __nv_bfloat16 *__restrict__ out) {
  for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < total; i += (long) gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn((float) w[i]);
}

// output neuron: omega = T - y, E += omega^2 / 2, psi_out = omega * f'(theta_out)   (network.cpp:242-243, 250)
__global__ void __launch_bounds__(256) out_delta_kernel(const double *__restrict__ y, const double *__restrict__ t, long n,
                                                         float *__restrict__ psi_out, double *__restrict__ err_sum) {
  double e = 0.0;
  for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
    const double yv = y[i], om = t[i] - yv;
    e += om * om / 2.0;
    psi_out[i] = (float) (om * (1.0 - yv * yv / 16.0));
  }
  for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
  if ((threadIdx.x & 31) == 0 && e != 0.0) atomicAdd(err_sum, e);
}

// dW of the last layer (N = 1): g[k] += sum_b psi_out[b] * a[b][k]. Thread = two columns, blockIdx.y = board slice.
__global__ void __launch_bounds__(256) gemv_t_kernel(const __nv_bfloat16 *__restrict__ a, const float *__restrict__ psi_out, long n_boards,
                                                      int K, long slice, double *__restrict__ g) {
  const int k2 = blockIdx.x * blockDim.x + threadIdx.x;
  if (k2 >= K / 2) return;
  const long b0 = blockIdx.y * slice, b1 = min(b0 + slice, n_boards);
  float s0 = 0.f, s1 = 0.f;
  for (long b = b0; b < b1; ++b) {
    const float2 v = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162 *>(a + b * K)[k2]);
    const float p = psi_out[b];
    s0 = fmaf(p, v.x, s0);
    s1 = fmaf(p, v.y, s1);
  }
  atomicAdd(g + 2 * k2, (double) s0);
  atomicAdd(g + 2 * k2 + 1, (double) s1);
}

// psi of the last hidden layer: psi[b][k] = psi_out[b] * w_last[k] * f'(a[b][k]), row-major and transposed.
// One CTA = a 64-board x 64-feature tile, transposed through shared memory.
__global__ void __launch_bounds__(256) psi_last_kernel(const __nv_bfloat16 *__restrict__ a, const float *__restrict__ w_last,
                                                        const float *__restrict__ psi_out, long n_boards, int K,
                                                        __nv_bfloat16 *__restrict__ psi, __nv_bfloat16 *__restrict__ psi_t, long ldt) {
  __shared__ __align__(16) __nv_bfloat16 tile[64][72];
  const long b0 = blockIdx.x * 64L;
  const int k0 = blockIdx.y * 64;
  {
    const int r = threadIdx.x >> 2, cq = (threadIdx.x & 3) * 16;
    const long b = b0 + r;
    uint4 raw[2] = {};
    float po = 0.f;
    if (b < n_boards) {
      const uint4 *src = reinterpret_cast<const uint4 *>(a + b * K + k0 + cq);
      raw[0] = __ldg(src); raw[1] = __ldg(src + 1);
      po = psi_out[b];
    }
    const __nv_bfloat162 *ab = reinterpret_cast<const __nv_bfloat162 *>(raw);
    __nv_bfloat162 out[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 v = __bfloat1622float2(ab[i]);
      const float p0 = po * w_last[k0 + cq + 2 * i] * (1.0f - 0.0625f * v.x * v.x);
      const float p1 = po * w_last[k0 + cq + 2 * i + 1] * (1.0f - 0.0625f * v.y * v.y);
      out[i] = __floats2bfloat162_rn(p0, p1);
      tile[cq + 2 * i][r] = out[i].x;
      tile[cq + 2 * i + 1][r] = out[i].y;
    }
    if (b < n_boards) {
      uint4 *dst = reinterpret_cast<uint4 *>(psi + b * K + k0 + cq);
      dst[0] = reinterpret_cast<const uint4 *>(out)[0];
      dst[1] = reinterpret_cast<const uint4 *>(out)[1];
    }
  }
  __syncthreads();
  {
    const int kk = threadIdx.x >> 2, bq = (threadIdx.x & 3) * 16;
    uint4 *dst = reinterpret_cast<uint4 *>(psi_t + (long) (k0 + kk) * ldt + b0 + bq);   // pad boards carry zeros
    dst[0] = *reinterpret_cast<const uint4 *>(&tile[kk][bq]);
    dst[1] = *reinterpret_cast<const uint4 *>(&tile[kk][bq + 8]);
  }
}

}  // namespace tc
}  // namespace cab
