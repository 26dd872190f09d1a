// mlp_tc_fused.cuh — the whole 4-layer MLP of Network::predictPosition (/root/reference/src/mcts_network/network.cpp:
// 92-117) as ONE tcgen05 kernel at REDUCED precision (fp16 operands, fp32 accumulation): the optional fast path of
// BASELINE.json's north_star for nets shaped {64, N1, N2, N3, 1} with N1, N2, N3 <= 256 (the default 64-144-225-100-1).
// Not a parity path: see cab_eval_f16_* in include/chessai_b200.h for the stated tolerance.
//
// One CTA per SM, persistent over tiles of 128 boards; TMEM lane t = board t of the tile.
//   * ALL weights live in shared memory for the whole kernel (fp16, zero-padded, pre-swizzled by f16_pack_kernel into
//     the K-major SWIZZLE_128B image tcgen05 reads through shared-memory descriptors): 167 KB for the default net,
//     fetched once per CTA with TMA bulk copies
//   * activations never touch shared or global memory (the board itself is staged in shared memory as layer 0's A operand):
//     the A operand of layers 1 and 2 is read FROM TENSOR MEMORY
//     (tcgen05.mma ... [d], [a], b_desc: SASS UTCHMMA with a TMEM A operand). A layer's accumulator row is loaded by its
//     thread (tcgen05.ld), put through the activation (carried as tanh(x/4), see act_pack_f32), packed to fp16 and
//     stored back IN PLACE (tcgen05.st) as the next layer's A row — lane t only ever touches lane t: no shuffle, no layout change
//   * the last layer (N = 1) is a dot product inside the threads that already hold the row
//   * epilogues and MMAs of consecutive layers and tiles overlap: see "the pipelined kernel" below
#pragma once
#include <cuda_fp16.h>

#include "cab_internal.cuh"
#include "tc_gemm_bf16.cuh"

namespace cab {
namespace tcf {

constexpr int kTile = 128;               // boards per tile = TMEM lanes
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
  long long *prof;            // optional [grid][64] clock stamps of the CTA's second tile: epilogue warp 0 [0, 16), MMA thread [16, 32),
                              // epilogue 1 end / start per warp [32, 48) / [48, 64) (CAB_F16_PROFILE debug launches)
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

// ---- the pipelined kernel ---------------------------------------------------------------------------------------------
// Sixteen epilogue warps (warps q, q + 4, q + 8, q + 12 serve TMEM lanes 32 q .. 32 q + 31: thread = board, the four
// warps of a lane quarter take every fourth chunk of a row) and one MMA thread run as a pipeline inside a tile and across
// tiles. The 512 TMEM columns are two 256-column regions P and Q that alternate between "accumulator" and "A operand":
//   A0 (the board as fp16, K = 64)  SHARED MEMORY, two buffers (K-major SWIZZLE_128B like the weights): layer 0 is an
//                                   "SS" MMA, so it needs no tensor-memory columns and can be issued a tile ahead
//   layer 0   A0 -> D0 = Q[0, n0)                     epilogue 0 packs f(D0) IN PLACE into Q[0, n0 / 2) = A1
//   layer 1   A1 -> D1 = P[0, n1)                     epilogue 1 packs f(D1) in place into P[0, n1 / 2) = A2
//   layer 2   A2 -> D2 = Q[n0, n0 + n2)               final epilogue: f(D2) . w_last, y = f(sum)
// In-place packing: the CH / 2 columns written for chunk c lie inside chunk c / 2, which has been READ before - by the
// same warp in program order, or by another warp of the quarter, which publishes the chunks it has read in a shared-
// memory counter. Four K steps of the next layer (64 columns) form a group: the epilogue warps arrive on
// e_ready[layer][group] per chunk and the MMA thread issues the group at once, so the tensor pipe works under the
// epilogue instead of after it.
// Order of the MMA thread: L1(t), L0(t + 1), L2(t): D0 of the next tile is complete long before its epilogue starts.
// Order of the epilogue warps: board(t + 1) -> epilogue 0(t) -> FINAL(t - 1) -> epilogue 1(t): the final epilogue of a
// tile is deferred into the next one, where the warps would otherwise wait for the tail of layer 1's MMAs.
// Every barrier completes exactly once per tile (a0_ready: once per two tiles): its parity follows the tile counter.
// Measured floors per 128-board tile (tools/mufu_bench.cu, tmem_bench.cu, mma_bench.cu): 496 MUFU.TANH per board thread at
// 8 clocks per warp instruction and SM sub-partition = 4.0 K clocks; MMAs 2.2 K clocks when issued in unrolled groups.
// (n0 + n2 > 256, possible for other shapes than the default net: D2 = Q[0, n2), nothing is deferred or issued ahead.)
constexpr int kPerQuarter = 4;             // epilogue warps per TMEM lane quarter (warps q, q + 4, q + 8, q + 12 all reach quarter q)
constexpr int kEpiWarps = 4 * kPerQuarter;
constexpr int kThreadsP = (kEpiWarps + 1) * 32;
constexpr int kColP = 0, kColQ = 256;
constexpr uint32_t kBoardBytes = kTile * 128;   // one A0 buffer: 128 rows of 64 fp16
constexpr int kMaxGroups = 4;              // 64-column groups (K blocks of the next layer) of a 256-wide layer

__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// one warp's arrival on a barrier the MMA thread waits on: its tensor-memory stores are complete and ordered first
__device__ __forceinline__ void warp_publish(uint64_t *bar, int lane) {
  tmem_wait_st();
  tc::tmem_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// The activation on two accumulator values, result packed as the fp16 pair the next layer reads.
// f(x) = 4 tanh(x / 4), but neither factor 4 is computed here: the kernel carries t = tanh(x / 4) = a / 4 between the
// layers. A layer fed with t instead of a = 4 t and asked for x / 4 instead of x needs its weights times 4 / 4 = 1, so
// only the FIRST layer's image is scaled (by 1/4, exact: its input is k * 0.4, not a tanh) and the output is 4 * tanh of
// the last dot product. Two fp32 hardware tanh and one pack to fp16x2 (tanh.approx.f16x2 is two MUFU.TANH.F16 plus a
// permute in SASS: the same MUFU count, one instruction more and a less accurate result).
__device__ __forceinline__ uint32_t act_pack_f32(float a, float b) {
  float ta, tb;
  asm("tanh.approx.f32 %0, %1;" : "=f"(ta) : "f"(a));
  asm("tanh.approx.f32 %0, %1;" : "=f"(tb) : "f"(b));
  return pack_h2(ta, tb);
}

template <int CH>
__device__ __forceinline__ void ld_chunk(uint32_t row_base, int col0, int n, uint32_t (&r)[CH / 16][16]) {
#pragma unroll
  for (int j = 0; j < CH / 16; ++j)   // unconditional (a load under a predicate costs a register copy per value): past the row's end
    tmem_ld16_async(row_base + min(col0 + 16 * j, n - 16), r[j]);   // the last 16 columns are simply read again and not used
}
template <int CH>
__device__ __forceinline__ void act_chunk(const uint32_t (&r)[CH / 16][16], uint32_t (&packed)[CH / 16][8]) {
#pragma unroll
  for (int j = 0; j < CH / 16; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) packed[j][i] = act_pack_f32(__uint_as_float(r[j][2 * i]), __uint_as_float(r[j][2 * i + 1]));
}
template <int CH>
__device__ __forceinline__ void st_chunk(uint32_t row_base, int col0, int n, const uint32_t (&packed)[CH / 16][8]) {
#pragma unroll
  for (int j = 0; j < CH / 16; ++j)
    if (col0 + 16 * j < n) tmem_st8(row_base + (col0 + 16 * j) / 2, packed[j]);
}

// The warps of a lane quarter tell each other how far they have READ: progress[warp] = 16 * (epilogue instance) + 1 +
// the last chunk that warp has read (it reads its chunks c = h, h + kPerQuarter, ... in ascending order).
struct QuarterProgress {
  volatile int *quarter;                      // progress[] of this lane quarter's warps: quarter[4 * h']
  int base;                                   // 16 * instance
  int h;                                      // which chunks are mine: c % kPerQuarter == h
  __device__ __forceinline__ void read_done(int c, int lane) const {   // after tcgen05.wait::ld of chunk c (warp-synchronous)
    if (lane == 0) quarter[4 * h] = base + c + 1;
  }
  __device__ __forceinline__ void before_store(int c) const {          // the columns chunk c is packed into belong to chunk c / 2
    const int j = c >> 1, owner = j % kPerQuarter;
    if (owner != h) while (quarter[4 * owner] < base + j + 1) {}
  }
};

// f() on one accumulator row of n columns, this warp's chunks (c = h, h + kPerQuarter, ...), packed in place as the next
// layer's A row. The MMA thread issues the next layer in GROUPS of four K steps (64 columns = one K block of the weight
// image): a chunk arrives on the barrier of its group, e_ready[64-column group], whose count is the group's chunks x 4.
template <int CH>
__device__ __forceinline__ void epilogue_pack(uint32_t row_base, int n, uint64_t *e_ready, const QuarterProgress pp, int lane) {
  uint32_t r[CH / 16][16], packed[CH / 16][8];
  const int nc = (n + CH - 1) / CH;
#pragma unroll 1
  for (int c = pp.h; c < nc; c += kPerQuarter) {   // the other warps of the quarter fill this warp's load and store latencies
    ld_chunk<CH>(row_base, c * CH, n, r);
    tmem_wait_ld();
    pp.read_done(c, lane);
    act_chunk<CH>(r, packed);
    pp.before_store(c);
    st_chunk<CH>(row_base, c * CH, n, packed);
    warp_publish(&e_ready[(c * CH) >> 6], lane);
  }
}

// the last hidden layer's row: f() and the N = 1 output layer as a dot product (fp32); w_last is a shared-memory address
template <int CH>
__device__ __forceinline__ float dot_chunk(const uint32_t (&r)[CH / 16][16], int col0, int n, uint32_t w_last, float acc) {
  float d0 = acc, d1 = 0.0f;
#pragma unroll
  for (int j = 0; j < CH / 16; ++j)
    if (col0 + 16 * j < n) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 w;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(w.x), "=f"(w.y), "=f"(w.z), "=f"(w.w) : "r"(w_last + (uint32_t) (col0 + 16 * j + 4 * i) * 4u));
        float t0, t1, t2, t3;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(__uint_as_float(r[j][4 * i])));
        asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(__uint_as_float(r[j][4 * i + 1])));
        asm("tanh.approx.f32 %0, %1;" : "=f"(t2) : "f"(__uint_as_float(r[j][4 * i + 2])));
        asm("tanh.approx.f32 %0, %1;" : "=f"(t3) : "f"(__uint_as_float(r[j][4 * i + 3])));
        d0 = fmaf(t0, w.x, d0);
        d1 = fmaf(t1, w.y, d1);
        d0 = fmaf(t2, w.z, d0);
        d1 = fmaf(t3, w.w, d1);
      }
    }
  return d0 + d1;
}
template <int CH>
__device__ __forceinline__ float epilogue_dot(uint32_t row_base, int n, uint32_t w_last, int h) {
  uint32_t r[CH / 16][16];
  const int nc = (n + CH - 1) / CH;
  float dot = 0.0f;
#pragma unroll 1
  for (int c = h; c < nc; c += kPerQuarter) {
    ld_chunk<CH>(row_base, c * CH, n, r);
    tmem_wait_ld();
    dot = dot_chunk<CH>(r, c * CH, n, w_last, dot);
  }
  return dot;
}

template <int CH>
__global__ void __launch_bounds__(kThreadsP, 1) mlp_tc_fused_kernel(const FusedArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
  uint8_t *a0_buf = smem + ((args.image_bytes + 1023u) & ~1023u);               // [2][128 rows][128 B], swizzle atoms of 1 KB
  uint64_t *bars = reinterpret_cast<uint64_t *>(a0_buf + 2 * kBoardBytes);
  uint64_t *bar_w = bars;                    // [3] a layer's weight image has landed (the last one brings w_last along)
  uint64_t *a0_ready = bars + 3;             // [2] a tile's boards are in shared memory               (epilogue warps -> MMA)
  uint64_t *d_full = bars + 5;               // [3] a layer's MMAs have retired                        (MMA -> epilogue warps)
  uint64_t *d2_free = bars + 8;              // the final epilogue has read D2                         (epilogue warps -> MMA)
  uint64_t *e_ready = bars + 9;              // [2][kMaxGroups] a 64-column group of A1 / A2 is in tensor memory (epilogue warps -> MMA)
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 9 + 2 * kMaxGroups);
  volatile int *progress = reinterpret_cast<volatile int *>(tmem_slot + 2);     // [kEpiWarps] chunks read, per epilogue warp
  float *dot_part = reinterpret_cast<float *>(tmem_slot + 2 + kEpiWarps);       // [2][kPerQuarter - 1][128] the other warps' shares of the dot product
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int l = 0; l < 3; ++l) mbar_init(&bar_w[l], 1);
    mbar_init(&a0_ready[0], kEpiWarps);
    mbar_init(&a0_ready[1], kEpiWarps);
    for (int l = 0; l < 3; ++l) mbar_init(&d_full[l], 1);
    mbar_init(d2_free, kEpiWarps);
    for (int l = 0; l < 2; ++l)                // e_ready[l][g]: every chunk of 64-column group g, from each of the four lane quarters
      for (int g = 0; g < kMaxGroups; ++g) {
        const int cols = min(64, max(args.n_pad[l] - 64 * g, 0));
        mbar_init(&e_ready[l * kMaxGroups + g], 4 * max((cols + CH - 1) / CH, 1));
      }
    for (int i = 0; i < kEpiWarps; ++i) progress[i] = 0;
    mbar_fence_init();
    // the image in three pieces, layer 0 first: a CTA of a small launch lives for one tile, and layer 0 (18 KB of the
    // default net's 167 KB) can start while the rest is still on its way
    for (int l = 0; l < 3; ++l) {
      const uint32_t lo = args.w_off[l], hi = l < 2 ? args.w_off[l + 1] : args.image_bytes;
      mbar_arrive_expect_tx(&bar_w[l], hi - lo);
      for (uint32_t off = lo; off < hi; off += 32768) bulk_g2s(smem + off, args.image + off, min(32768u, hi - off), &bar_w[l]);
    }
  }
  if (warp == kEpiWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc::tmem_fence_before();
  __syncthreads();
  tc::tmem_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long n_tiles = (args.n_boards + kTile - 1) / kTile;
  const int n0 = args.n_pad[0], n1 = args.n_pad[1], n2 = args.n_pad[2];
  const int nk1 = n0 >> 4, nk2 = n1 >> 4;                         // K steps of layer 1 / 2 (= the previous padded width / 16)
  const bool overlap_final = n0 + n2 <= 256;
  const uint32_t col_d2 = kColQ + (overlap_final ? (uint32_t) n0 : 0u);

  if (warp == kEpiWarps) {
    // ---- the MMA thread ----
    // Four K steps (one 64-wide K block of the weight image) per barrier, unrolled, their descriptors a base plus constants:
    // every MMA in flight then reads its own uniform registers. Issued one by one from a rolled loop with the descriptor
    // recomputed per step, a tcgen05.mma costs ~270 clocks whatever its shape (tools/mma_bench.cu) instead of 128 N / 256.
    if (lane == 0) {
      mbar_wait(&bar_w[0], 0);
      const uint32_t idesc0 = idesc_f16(n0), idesc1 = idesc_f16(n1), idesc2 = idesc_f16(n2);
      const uint32_t w0 = smem_u32(smem + args.w_off[0]), w1 = smem_u32(smem + args.w_off[1]), w2 = smem_u32(smem + args.w_off[2]);
      // K block g of a layer: B = np rows x 128 bytes, 32 bytes further per K step inside the swizzle atom; A = 8 columns per step
      auto issue_group = [&](uint32_t d_col, uint32_t a_col, uint32_t w_base, int np, uint32_t idesc, int g, int nk) {
        const uint64_t b = tc::make_smem_desc(w_base + (uint32_t) g * (uint32_t) np * 128u);
        const uint32_t a = tmem_base + a_col + (uint32_t) g * 32u, d = tmem_base + d_col;
        const int left = nk - 4 * g;           // straight-line code per group size: no branch between two MMAs
        if (left >= 4) {
          umma_ts(d, a, b, idesc, g != 0 ? 1u : 0u);
          umma_ts(d, a + 8u, b + 2, idesc, 1u);
          umma_ts(d, a + 16u, b + 4, idesc, 1u);
          umma_ts(d, a + 24u, b + 6, idesc, 1u);
        } else if (left == 3) {
          umma_ts(d, a, b, idesc, g != 0 ? 1u : 0u);
          umma_ts(d, a + 8u, b + 2, idesc, 1u);
          umma_ts(d, a + 16u, b + 4, idesc, 1u);
        } else if (left == 2) {
          umma_ts(d, a, b, idesc, g != 0 ? 1u : 0u);
          umma_ts(d, a + 8u, b + 2, idesc, 1u);
        } else if (left == 1) {
          umma_ts(d, a, b, idesc, g != 0 ? 1u : 0u);
        }
      };
      // layer 0 of the it-th tile: A from the shared-memory board buffer it & 1 ("SS"), 32 bytes further per K step
      auto issue_layer0 = [&](uint32_t it0) {
        mbar_wait(&a0_ready[it0 & 1], (it0 >> 1) & 1);
        tc::tmem_fence_after();
        const uint64_t a = tc::make_smem_desc(smem_u32(a0_buf + (it0 & 1) * kBoardBytes)), b = tc::make_smem_desc(w0);
        const uint32_t d = tmem_base + kColQ;
        tc::umma_f16(d, a, b, idesc0, 0u);                 // K = 64: four steps (k_steps[0] == 4)
        tc::umma_f16(d, a + 2, b + 2, idesc0, 1u);
        tc::umma_f16(d, a + 4, b + 4, idesc0, 1u);
        tc::umma_f16(d, a + 6, b + 6, idesc0, 1u);
        tc::umma_commit(&d_full[0]);
      };
      uint32_t it = 0;
      if ((long) blockIdx.x < n_tiles) issue_layer0(0);
      for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t par = it & 1;
        const bool has_next = tile + gridDim.x < n_tiles;
        long long *stamp = (args.prof != nullptr && it == 1) ? args.prof + (size_t) blockIdx.x * 64 + 16 : nullptr;
        if (stamp) stamp[0] = clock64();
        if (it == 0) mbar_wait(&bar_w[1], 0);
        for (int g = 0; 4 * g < nk1; ++g) {
          mbar_wait(&e_ready[g], par);
          tc::tmem_fence_after();
          if (stamp && g == 0) stamp[3] = clock64();
          issue_group(kColP, kColQ, w1, n1, idesc1, g, nk1);
        }
        if (stamp) stamp[4] = clock64();
        tc::umma_commit(&d_full[1]);
        if (stamp) stamp[5] = clock64();
        // D0's columns are free as soon as layer 1 has been issued (the tensor pipe runs in order): the next tile's layer 0
        if (overlap_final && has_next) issue_layer0(it + 1);
        if (stamp) stamp[2] = clock64();
        long long waited = 0;
        if (it == 0) mbar_wait(&bar_w[2], 0);
        for (int g = 0; 4 * g < nk2; ++g) {
          const long long w0c = stamp ? clock64() : 0;
          mbar_wait(&e_ready[kMaxGroups + g], par);
          if (stamp) waited += clock64() - w0c;
          if (g == 0 && overlap_final && it > 0) mbar_wait(d2_free, (it - 1) & 1);   // the previous tile's D2 has been read
          tc::tmem_fence_after();
          if (stamp && g == 0) stamp[6] = clock64();
          issue_group(col_d2, kColP, w2, n2, idesc2, g, nk2);
        }
        if (stamp) stamp[7] = clock64();
        tc::umma_commit(&d_full[2]);
        if (stamp) { stamp[8] = clock64(); stamp[9] = stamp[0] + waited; stamp[1] = stamp[0]; }
        if (!overlap_final && has_next) {      // D2 shares D0's columns: the next layer 0 waits for this tile's final epilogue
          mbar_wait(d2_free, par);
          issue_layer0(it + 1);
        }
      }
    }
    __syncwarp();
  } else {
    // ---- the epilogue warps: thread = board, warp h of a lane quarter = the chunks c with c % kPerQuarter == h ----
    const int quarter = warp & 3, h = warp >> 2;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t) (quarter * 32) << 16);
    const uint32_t w_last = smem_u32(smem + args.w_last_off);
    QuarterProgress pp;
    pp.quarter = progress + quarter;
    pp.h = h;
    constexpr int kCodeVecs = 4 / kPerQuarter;            // uint4 (16 codes) per warp: the board is split over the quarter's warps
    static_assert(kPerQuarter == 1 || kPerQuarter == 2 || kPerQuarter == 4, "the 64 codes are split in whole 16-byte pieces");
    uint4 cd[kCodeVecs];
    auto fetch_codes = [&](long tile) {
      const long b = tile * kTile + row_in_tile;
#pragma unroll
      for (int i = 0; i < kCodeVecs; ++i) cd[i] = make_uint4(0, 0, 0, 0);
      if (tile < n_tiles && b < args.n_boards) {
        const uint4 *src = reinterpret_cast<const uint4 *>(args.codes + b * 64) + kCodeVecs * h;
#pragma unroll
        for (int i = 0; i < kCodeVecs; ++i) cd[i] = __ldg(src + i);
      }
    };
    // Network::loadBoard: codes -> fp16 k * 0.4 -> this board's row of the A0 buffer `buf` (64 fp16 = one 128-byte row of
    // a K-major SWIZZLE_128B tile: 16-byte piece c of row r sits at piece c ^ (r % 8); 16 codes = two pieces per uint4)
    auto store_board = [&](uint32_t buf) {
      uint8_t *row = a0_buf + buf * kBoardBytes + (row_in_tile >> 3) * 1024 + (row_in_tile & 7) * 128;
#pragma unroll
      for (int p2 = 0; p2 < kCodeVecs; ++p2) {
        const uint32_t w4[4] = {cd[p2].x, cd[p2].y, cd[p2].z, cd[p2].w};
        uint32_t packed[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t word = w4[i >> 1] >> ((i & 1) * 16);
          packed[i] = pack_h2((float) (int) (int8_t) (word & 0xff) * 0.4f, (float) (int) (int8_t) ((word >> 8) & 0xff) * 0.4f);
        }
        const int piece = 2 * (kCodeVecs * h + p2);
        *reinterpret_cast<uint4 *>(row + ((piece ^ (row_in_tile & 7)) << 4)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        *reinterpret_cast<uint4 *>(row + (((piece + 1) ^ (row_in_tile & 7)) << 4)) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&a0_ready[buf]);
    };
    fetch_codes(blockIdx.x);
    if ((long) blockIdx.x < n_tiles) store_board(0);
    fetch_codes((long) blockIdx.x + gridDim.x);
    // the last hidden layer of tile `tile_f` (the it_f-th of this CTA): f(D2) . w_last, y = f(sum)
    auto final_epilogue = [&](long tile_f, uint32_t it_f, long long *stamp) {
      const uint32_t par_f = it_f & 1;
      const long board = tile_f * kTile + row_in_tile;
      mbar_wait(&bar_w[2], 0);                 // w_last arrives with the last piece of the image (a completed phase: free afterwards)
      mbar_wait(&d_full[2], par_f);
      tc::tmem_fence_after();
      if (stamp) stamp[6] = clock64();
      const float dot = epilogue_dot<CH>(lane_base + col_d2, n2, w_last, h);
      if (stamp) stamp[7] = clock64();
      tc::tmem_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d2_free);
      float *part = dot_part + par_f * (kPerQuarter - 1) * kTile;
      if (h > 0) part[(h - 1) * kTile + row_in_tile] = dot;
      asm volatile("bar.sync %0, %1;" ::"r"(quarter + 1), "n"(kPerQuarter * 32) : "memory");   // the warps of this lane quarter
      if (h == 0 && board < args.n_boards) {
        float sum = dot;
#pragma unroll
        for (int i = 0; i < kPerQuarter - 1; ++i) sum += part[i * kTile + row_in_tile];
        float y;
        asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(sum));   // the sum already is x / 4
        args.out[board] = (double) (4.0f * y);
      }
    };
    // With D2 beside D0 the final epilogue of a tile is DEFERRED into the next tile, between its epilogues 0 and 1: there
    // the warps would otherwise wait for the tail of layer 1's MMAs, and the wait for layer 2's own tail disappears.
    const bool defer_final = overlap_final;
    uint32_t it = 0;
    long prev_tile = -1;
    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      const long next = tile + gridDim.x;
      // the next tile's boards (fetched a tile ago) go to the other buffer: its last reader, layer 0 of tile t - 1, retired
      // before epilogue 0 of that tile began. Then the codes of the tile after are requested, a whole tile ahead of use.
      if (next < n_tiles) store_board((it + 1) & 1);
      fetch_codes(next + gridDim.x);
      long long *stamp = (args.prof != nullptr && it == 1 && warp == 0 && lane == 0) ? args.prof + (size_t) blockIdx.x * 64 : nullptr;
      if (stamp) stamp[0] = clock64();
#pragma unroll 1
      for (int l = 0; l < 2; ++l) {            // epilogue 0: D0 in Q -> A1; epilogue 1: D1 in P -> A2
        mbar_wait(&d_full[l], par);
        tc::tmem_fence_after();
        if (stamp) stamp[1 + 2 * l] = clock64();
        if (args.prof != nullptr && it == 1 && l == 1 && lane == 0) args.prof[(size_t) blockIdx.x * 64 + 48 + warp] = clock64();
        pp.base = (int) (it * 2 + (uint32_t) l + 1) * 16;
        epilogue_pack<CH>(lane_base + (l == 0 ? kColQ : kColP), l == 0 ? n0 : n1, e_ready + l * kMaxGroups, pp, lane);
        if (stamp) stamp[2 + 2 * l] = clock64();
        if (args.prof != nullptr && it == 1 && l == 1 && lane == 0) args.prof[(size_t) blockIdx.x * 64 + 32 + warp] = clock64();
        if (l == 0 && defer_final && prev_tile >= 0) final_epilogue(prev_tile, it - 1, stamp);
      }
      if (stamp) stamp[5] = clock64();
      if (!defer_final) final_epilogue(tile, it, stamp);
      prev_tile = tile;
    }
    if (defer_final && prev_tile >= 0) final_epilogue(prev_tile, it - 1, nullptr);
  }

  tc::tmem_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) {
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
      if (l == 0) v *= 0.25;      // the kernel carries tanh(x / 4) between layers (see act_pack_f32): only layer 0 is rescaled
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
