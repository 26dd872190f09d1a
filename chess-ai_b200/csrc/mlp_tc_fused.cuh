// mlp_tc_fused.cuh — the whole 4-layer MLP of Network::predictPosition (/root/reference/src/mcts_network/network.cpp:
// 92-117) as ONE tcgen05 kernel at REDUCED precision (fp16 operands, fp32 accumulation): the optional fast path of
// BASELINE.json's north_star for nets shaped {64, N1, N2, N3, 1} with N1, N2, N3 <= 256 (the default 64-144-225-100-1).
// Not a parity path: see cab_eval_f16_* in include/chessai_b200.h for the stated tolerance.
//
// One CTA per SM, persistent over tiles of 128 boards; TMEM lane t = board t of the tile. Two threads serve a lane
// (warps w and w + 4 both reach lane quarter w % 4): they split the 16-column chunks of every epilogue between them.
//   * ALL weights live in shared memory for the whole kernel (fp16, zero-padded, pre-swizzled by f16_pack_kernel into
//     the K-major SWIZZLE_128B image tcgen05 reads through shared-memory descriptors): 167 KB for the default net,
//     fetched once per CTA with TMA bulk copies
//   * activations never touch shared or global memory: the A operand of every layer is read FROM TENSOR MEMORY
//     (tcgen05.mma ... [d], [a], b_desc: SASS UTCHMMA with a TMEM A operand). A layer's accumulator row is loaded by its
//     thread (tcgen05.ld), put through the activation (carried as tanh(x/4), see act_pack_h2), packed to fp16 and
//     stored straight back (tcgen05.st) as the next layer's A row — lane t only ever touches lane t, so there is no shuffle and no layout change
//   * TMEM columns: [0, 128) the A operand (up to 256 fp16 per row), [128, 384) the fp32 accumulator
//   * the last layer (N = 1) is a dot product inside the thread that already holds the row
#pragma once
#include <cuda_fp16.h>

#include "cab_internal.cuh"
#include "tc_gemm_bf16.cuh"

namespace cab {
namespace tcf {

constexpr int kTile = 128;               // boards per tile = TMEM lanes
constexpr int kThreads = 256;
constexpr int kColA = 0, kColD = 128;    // TMEM column bases
constexpr int kTmemCols = 512;
constexpr int kMaxLayers = 3;            // tensor-core layers (the 4th, N = 1, is a dot product)

struct FusedArgs {
  const uint8_t *image;       // packed fp16 weight image (all tensor-core layers) + fp32 last-layer vector
  uint32_t image_bytes;       // multiple of 16
  uint32_t w_off[kMaxLayers]; // byte offset of each layer's image
  uint32_t w_last_off;        // byte offset of the fp32 last-layer weights (n_pad[2] floats)
  int k_pad[kMaxLayers];      // contraction length of the weight image, multiple of 64 (whole swizzle atoms)
  int k_steps[kMaxLayers];    // K steps of 16 the MMAs actually run: ceil(real K / 16)
  int n_pad[kMaxLayers];      // outputs, multiple of 16
  const int8_t *codes;        // [n_boards][64]
  long n_boards;
  double *out;                // [n_boards]
};

// kind::f16 instruction descriptor: D fp32 (bit 4), A and B fp16 (formats 0), K-major both, N >> 3 at bit 17, M = 128
__device__ __forceinline__ uint32_t idesc_f16(int n) { return (1u << 4) | ((uint32_t) (n >> 3) << 17) | ((uint32_t) (kTile >> 4) << 24); }

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t *>(&h);
}

// The activation on two accumulator values at once, result already packed as the fp16 pair the next layer reads.
// f(x) = 4 tanh(x / 4), but neither factor 4 is computed here: the kernel carries t = tanh(x / 4) = a / 4 between the
// layers. A layer fed with t instead of a = 4 t and asked for x / 4 instead of x needs its weights times 4 / 4 = 1, so
// only the FIRST layer's image is scaled (by 1/4, exact: its input is k * 0.4, not a tanh) and the output is 4 * tanh of
// the last dot product. The epilogue is then one cvt.rn.f16x2.f32 and ONE tanh.approx.f16x2 per pair of values.
__device__ __forceinline__ uint32_t act_pack_h2(float a, float b) {
  const __half2 q = __floats2half2_rn(a, b);
  uint32_t t;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(*reinterpret_cast<const uint32_t *>(&q)));
  return t;
}

__global__ void __launch_bounds__(kThreads, 1) mlp_tc_fused_kernel(const FusedArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
  uint64_t *bar_w = reinterpret_cast<uint64_t *>(smem + args.image_bytes);   // weights landed
  uint64_t *bar_mma = bar_w + 1;                                             // a layer's MMAs retired
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_mma + 1);
  float *dot_half = reinterpret_cast<float *>(tmem_slot + 4);                // [128] partial dot products of the upper half
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3, half = warp >> 2, row_in_tile = quarter * 32 + lane;

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(bar_w, args.image_bytes);
    for (uint32_t off = 0; off < args.image_bytes; off += 32768) {
      const uint32_t n = min(32768u, args.image_bytes - off);
      bulk_g2s(smem + off, args.image + off, n, bar_w);
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc::tmem_fence_before();
  __syncthreads();
  tc::tmem_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = tmem_base + ((uint32_t) (quarter * 32) << 16);   // this warp's TMEM lane quarter
  const float *w_last = reinterpret_cast<const float *>(smem + args.w_last_off);
  mbar_wait(bar_w, 0);

  uint32_t mma_phase = 0;
  const long n_tiles = (args.n_boards + kTile - 1) / kTile;
  // this thread's 32 codes of a board (its half of the 64), fetched one tile ahead: the global latency is off the path
  auto fetch_codes = [&](long tile, uint4 &r0, uint4 &r1) {
    const long b = tile * kTile + row_in_tile;
    r0 = r1 = make_uint4(0, 0, 0, 0);
    if (tile < n_tiles && b < args.n_boards) {
      const uint4 *src = reinterpret_cast<const uint4 *>(args.codes + b * 64) + 2 * half;
      r0 = __ldg(src);
      r1 = __ldg(src + 1);
    }
  };
  uint4 next0, next1;
  fetch_codes(blockIdx.x, next0, next1);
  for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long board = tile * kTile + row_in_tile;
    const bool valid = board < args.n_boards;
    // ---- Network::loadBoard: 64 codes -> fp16 k * 0.4 -> this lane's A row, K = 64 = 32 columns ----
    {
      const uint4 cur[2] = {next0, next1};
      fetch_codes(tile + gridDim.x, next0, next1);
#pragma unroll
      for (int p2 = 0; p2 < 2; ++p2) {
        const uint32_t w4[4] = {cur[p2].x, cur[p2].y, cur[p2].z, cur[p2].w};
        uint32_t packed[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t word = w4[i >> 1] >> ((i & 1) * 16);
          packed[i] = pack_h2((float) (int) (int8_t) (word & 0xff) * 0.4f, (float) (int) (int8_t) ((word >> 8) & 0xff) * 0.4f);
        }
        tmem_st8(lane_base + kColA + (2 * half + p2) * 8, packed);
      }
      tmem_wait_st();
    }
    float dot = 0.0f;
#pragma unroll 1
    for (int l = 0; l < kMaxLayers; ++l) {
      const int np = args.n_pad[l];
      tc::tmem_fence_before();
      __syncthreads();                               // every row of the A operand is in tensor memory
      if (threadIdx.x == 0) {
        tc::tmem_fence_after();
        const uint32_t idesc = idesc_f16(np);
        const uint32_t w_base = smem_u32(smem + args.w_off[l]);
        for (int k = 0; k < args.k_steps[l]; ++k) {   // only the K steps that hold data: the previous width padded to 16
          // B: K block k / 4 (np rows x 128 bytes each), 32 bytes further per K step inside the swizzle atom
          const uint64_t b_desc = tc::make_smem_desc(w_base + (uint32_t) (k >> 2) * (uint32_t) np * 128u) + (uint64_t) ((k & 3) * 2);
          umma_ts(tmem_base + kColD, tmem_base + kColA + (uint32_t) k * 8u, b_desc, idesc, k != 0 ? 1u : 0u);
        }
        tc::umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, mma_phase);
      mma_phase ^= 1;
      tc::tmem_fence_after();
      if (l + 1 < kMaxLayers) {
        // f() on this board's row, packed to fp16, stored back as the next layer's A row
#pragma unroll 1
        for (int c = 16 * half; c < np; c += 32) {
          uint32_t r[16], packed[8];
          tmem_ld16(lane_base + kColD + c, r);
#pragma unroll
          for (int i = 0; i < 8; ++i) packed[i] = act_pack_h2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
          tmem_st8(lane_base + kColA + c / 2, packed);
        }
        tmem_wait_st();                                // (n_pad is the next layer's K: no K padding left to zero)
      } else {
        // last hidden layer: keep fp32 and fold the N = 1 output layer in as a dot product
#pragma unroll 1
        for (int c = 16 * half; c < np; c += 32) {
          uint32_t r[16];
          tmem_ld16(lane_base + kColD + c, r);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t h = act_pack_h2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
            const float2 v = __half22float2(*reinterpret_cast<const __half2 *>(&h));
            dot = fmaf(v.x, w_last[c + 2 * i], dot);
            dot = fmaf(v.y, w_last[c + 2 * i + 1], dot);
          }
        }
      }
    }
    if (half == 1) dot_half[row_in_tile] = dot;
    tc::tmem_fence_before();
    __syncthreads();                                 // the accumulator has been read: the next tile may overwrite it
    if (half == 0 && valid) {
      float y;
      asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(dot + dot_half[row_in_tile]));   // the sum already is x / 4
      args.out[board] = (double) (4.0f * y);
    }
  }

  tc::tmem_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::tmem_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
}

// fp64 weights -> the shared-memory image. One thread per (layer, n, k) element of the padded matrices.
// Element (n, k) of layer l: K block kb = k / 64, 16-byte chunk c = (k % 64) / 8, byte offset
//   w_off[l] + kb * n_pad * 128 + (n / 8) * 1024 + (n % 8) * 128 + ((c ^ (n % 8)) * 16) + (k % 8) * 2      (SWIZZLE_128B, K-major)
struct PackArgs {
  const double *w;
  long woff[kMaxLayers + 1];
  int K[kMaxLayers + 1], N[kMaxLayers + 1];
  uint32_t w_off[kMaxLayers];
  uint32_t w_last_off;
  int k_pad[kMaxLayers], n_pad[kMaxLayers];
  uint8_t *image;
};
__global__ void __launch_bounds__(256) f16_pack_kernel(const PackArgs a) {
  for (int l = 0; l < kMaxLayers; ++l) {
    const long total = (long) a.n_pad[l] * a.k_pad[l];
    for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < total; i += (long) gridDim.x * blockDim.x) {
      const int n = (int) (i / a.k_pad[l]), k = (int) (i % a.k_pad[l]);
      double v = (n < a.N[l] && k < a.K[l]) ? a.w[a.woff[l] + (long) k * a.N[l] + n] : 0.0;   // W is [K][N]
      if (l == 0) v *= 0.25;      // the kernel carries tanh(x / 4) between layers (see act_pack_h2): only layer 0 is rescaled
      const uint32_t off = a.w_off[l] + (uint32_t) (k >> 6) * (uint32_t) a.n_pad[l] * 128u + (uint32_t) (n >> 3) * 1024u +
                           (uint32_t) (n & 7) * 128u + (uint32_t) ((((k & 63) >> 3) ^ (n & 7)) * 16) + (uint32_t) (k & 7) * 2u;
      *reinterpret_cast<__half *>(a.image + off) = __double2half(v);
    }
  }
  float *last = reinterpret_cast<float *>(a.image + a.w_last_off);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n_pad[kMaxLayers - 1]; i += gridDim.x * blockDim.x)
    last[i] = i < a.K[kMaxLayers] ? (float) a.w[a.woff[kMaxLayers] + i] : 0.0f;
}

}  // namespace tcf
}  // namespace cab
