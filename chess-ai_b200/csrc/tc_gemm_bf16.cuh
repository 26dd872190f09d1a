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

struct TcGemmArgs {
  int M, N, K;
  __nv_bfloat16 *Y;   // [M][N]
  int apply_act;
};

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
  const int n_tiles = n_blocks_m * n_blocks_n, n_kblocks = args.K / BK;

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
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int mb = tile / n_blocks_n, nb = tile % n_blocks_n;
        for (int kb = 0; kb < n_kblocks; ++kb) {
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
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[as], aph ^ 1);          // the epilogue drained this accumulator
      tmem_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t) as * BN;
      for (int kb = 0; kb < n_kblocks; ++kb) {
        mbar_wait(&full[s], ph);                    // TMA bytes landed
        tmem_fence_after();
        if (lane == 0) {
          const uint64_t a_desc = make_smem_desc(smem_u32(smem_a + (size_t) s * kBytesA));
          const uint64_t b_desc = make_smem_desc(smem_u32(smem_b + (size_t) s * kBytesB));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)     // advance 32 B inside the swizzle atom: +2 in the >>4 address field
            umma_f16(tmem_d, a_desc + (uint64_t) (k * 2), b_desc + (uint64_t) (k * 2), kIdesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty[s]);                   // frees the smem stage when these MMAs retire
          if (kb == n_kblocks - 1) umma_commit(&tmem_full[as]);
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
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int mb = tile / n_blocks_n, nb = tile % n_blocks_n;
      mbar_wait(&tmem_full[as], aph);
      tmem_fence_after();
      const int row = mb * BM + q * 32 + lane;
      __nv_bfloat16 *yrow = args.Y + (size_t) row * args.N + (size_t) nb * BN;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t) (q * 32) << 16) + (uint32_t) (as * BN + c), r);
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float v0 = __uint_as_float(r[2 * i]), v1 = __uint_as_float(r[2 * i + 1]);
          if (args.apply_act) { v0 = act_f32(v0); v1 = act_f32(v1); }
          __nv_bfloat162 b = __floats2bfloat162_rn(v0, v1);
          packed[i] = *reinterpret_cast<uint32_t *>(&b);
        }
        if (row < args.M) {
          uint4 *dst = reinterpret_cast<uint4 *>(yrow + c);
#pragma unroll
          for (int i = 0; i < 4; ++i) dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
        }
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

}  // namespace tc
}  // namespace cab
