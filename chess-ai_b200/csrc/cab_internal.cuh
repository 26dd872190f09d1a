// cab_internal.cuh — shared declarations for the CUDA side of libchessai_b200.so (not part of the ABI).
//
// Data layout in HBM:
//   w            flat fp64 weights, n,i,j order (the dump's line 3)                     n_weights * 8 B
//   pack_fwd     the same weights re-ordered into the DMMA B-fragment stream consumed by the forward kernel
//   pack_bwd     W^T in the same fragment order, consumed by the backward (omega) kernel
//   chunk tables descriptors of the <=16 KB pieces of those streams that one bulk copy moves to shared memory
// "unit"  = 8 consecutive features of 8 boards held as 32 lanes x double2 (512 B): lane (m = lane>>2, q = lane&3)
//           holds features 8t+2q, 8t+2q+1 of board m. It is simultaneously the DMMA C-fragment of an output tile
//           and the A-fragment pair of the next layer's two k-steps, so activations never change layout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/chessai_b200.h"

namespace cab {

constexpr int kConsumerWarps = 16;                      // max warps per CTA; a PAIR of warps owns 8 boards
constexpr int kMaxThreads = kConsumerWarps * 32;        // no producer warp: the last warp off a ring slot re-arms it
constexpr int kBoardsPerWarp = 8;                       // boards per warp pair (one DMMA M tile)
constexpr int kMaxStages = 8;                           // weight ring depth upper bound
constexpr int kStageBytes = 53248;                      // default ring slot (measured best: 2 x 52 KB with 8 pairs)
constexpr int kMaxNB = 32;                              // output tiles of one n-block (split over the pair)
constexpr int kMaxNBH = 16;                             // output tiles one warp accumulates in registers
constexpr int kUnitBytes = 512;                         // 32 lanes x double2

enum ChunkFlags : uint8_t { kFirst = 1, kLast = 2, kFinal = 4, kActIn = 8 };  // kActIn: apply f() to the A operand on load

struct __align__(16) ChunkDesc {
  uint32_t src_off;  // byte offset into the packed stream
  uint32_t bytes;    // multiple of 512
  uint16_t t0;       // first input unit covered by this chunk
  uint16_t tcount;   // input units in this chunk
  uint16_t tile0;    // first output tile of the n-block
  uint8_t nb;        // tiles in the n-block (1..kMaxNB)
  uint8_t flags;     // ChunkFlags
  uint16_t a_in;     // unit offset of the layer input inside the per-warp activation buffer
  uint16_t a_out;    // unit offset of this n-block's output
  uint16_t layer;    // weight layer index (0-based)
  uint16_t n_tiles;  // output tiles of the whole layer
  uint32_t save_off; // unit offset of this layer's output inside the saved-activation tensor (x n_boards_pad)
  uint32_t save_in;  // unit offset of this layer's INPUT activations inside the saved-activation tensor
};
static_assert(sizeof(ChunkDesc) == 32, "ChunkDesc must stay 32 bytes");

struct PackSeg {  // one (layer, n-block) piece of a packed stream, for the prepack kernel
  long dst_off;   // in doubles
  long w_off;     // offset of the layer's weights in the flat array
  int K, N;       // layer shape: W is [K][N]
  int t_in;       // input units
  int tile0, nb;  // output tiles covered
  int transposed; // 0: stream of W (forward); 1: stream of W^T (backward: contraction over N, outputs over K)
  long n_doubles;
};

struct StreamPlan {  // host description of one packed stream
  std::vector<ChunkDesc> chunks;
  std::vector<PackSeg> segs;
  long total_doubles = 0;
  int a_units = 0;      // per-warp activation buffer size in units
  ChunkDesc *d_chunks = nullptr;
  PackSeg *d_segs = nullptr;
  double *d_pack = nullptr;
};

constexpr int kCtxStreams = 4;  // H2D / kernel / D2H of consecutive chunks overlap across these
struct EvalCtx {  // per-caller scratch: lets many host threads evaluate on one net concurrently
  cudaStream_t stream[kCtxStreams] = {};
  int8_t *d_codes[kCtxStreams] = {};
  double *d_out[kCtxStreams] = {};
  uint8_t *d_records[kCtxStreams] = {};
  long cap_boards = 0;
  bool busy = false;
};

}  // namespace cab

struct cab_net {
  int device = 0;
  int sm_count = 148;
  int stage_bytes = cab::kStageBytes;  // ring slot size the stream plans were chunked for
  std::vector<int> dims;
  std::vector<long> woff;
  long n_weights = 0;
  double *d_w = nullptr;     // flat weights
  bool pack_dirty = true;    // packed streams stale w.r.t. d_w
  bool fp64_ok = true;       // the fp64 stream kernels fit this network's widest layer in shared memory
  bool tc_dirty = true;      // bf16 copies of the tcgen05 path stale w.r.t. d_w
  void *tc = nullptr;        // cab::TcState of the bf16 tcgen05 path (cab_tc.cu), created on first use
  cab::StreamPlan fwd, bwd;
  // saved-activation geometry (units per board-group, per layer)
  std::vector<uint32_t> act_unit_off;  // per layer l (0..L-1): first unit of layer l's activations
  uint32_t act_units_total = 0;
  // training state
  double *d_dw = nullptr;      // last applied update (connection_changes)
  double *d_grad = nullptr;    // n_weights + 4 doubles: sum over batch of a*psi, then [sumE, sumE', count, spare]
  double *d_acts = nullptr;    // saved activations [unit][board_pad][8]
  double *d_psis = nullptr;    // saved psi, same geometry
  long train_cap = 0;          // boards the two tensors above can hold
  int8_t *d_tcodes = nullptr;
  double *d_ttargets = nullptr;
  double *d_tout = nullptr;
  double *d_tout_tc = nullptr;  // outputs of the bf16 trainer's re-forward
  long tout_cap = 0;
  long tbuf_cap = 0;
  double *d_scalars = nullptr; // small device scratch (lambda, flags ...)
  cudaStream_t train_stream = nullptr;
  cudaStream_t aux_stream[4] = {};  // helper streams for fanned-out passes (cab_train.cu:fan_out)
  cudaEvent_t aux_event[5] = {};
  // data parallel
  void *nccl_comm = nullptr;
  int dp_rank = 0, dp_world = 1;
  // eval contexts
  std::mutex mu;
  std::vector<cab::EvalCtx *> ctxs;
  std::atomic<long> launches{0};
};

namespace cab {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define CAB_CUDA(call)                                                         \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) return cab::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)

// ---- device helpers -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar), ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// D(8x8) += A(8x4) * B(4x8) in fp64 on the tensor pipe (SASS: DMMA.8x8x4).
// lane: A[row = lane>>2][col = lane&3], B[row = lane&3][col = lane>>2], C[row = lane>>2][col = 2*(lane&3) + {0,1}]
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

// The reference activation, as written (mcts_network/network.h:83-94):
//   f(x) = MAX * modifiedSigmoid(x / WIDTH),  modifiedSigmoid(u) = 2/(1+exp(-2u)) - 1,  MAX = WIDTH = 4
__device__ __forceinline__ double act_f_libm(double x) { return 4.0 * (2.0 / (1.0 + exp(-2.0 * (x / 4.0))) - 1.0); }

// Same expression, same operation order, but exp and the division are branch-free so 8 of them pipeline per lane:
//   t = -2*(x/4) (exact), clamped to +-60: beyond that 1+e rounds to e or 1 and the reference's inf / 0 path and this one
//       both give exactly -+4
//   e = exp(t) = 2^n * P(r), n = rint(t*log2 e), r = t - n*ln2 (Cody-Waite), P = degree-13 Taylor (|r| <= 0.347:
//       truncation 4e-18 relative), scaled by adding n to the exponent field (|n| <= 87, result normal)
//   1/(1+e): MUFU.RCP64H seed + 3 Newton steps (error squares each step from ~2^-20)
//   f = 4*(2/(1+e) - 1) = fma(8, 1/(1+e), -4)  (the *4 is exact, so one rounding either way)
// Differences from libm-based evaluation are O(1 ulp) of the intermediate e, i.e. <= ~5e-16 absolute on f.
__device__ __forceinline__ double act_f(double x) {
  double t = -0.5 * x;
  t = fmin(fmax(t, -60.0), 60.0);
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: adding it rounds to nearest integer in the low bits
  const double nm = fma(t, 1.4426950408889634, magic);
  const double nd = nm - magic;
  double r = fma(nd, -6.93147180369123816490e-01, t);
  r = fma(nd, -1.90821492927058770002e-10, r);
  double p = 1.6059043836821613e-10;            // 1/13!
  p = fma(p, r, 2.08767569878681e-09);          // 1/12!
  p = fma(p, r, 2.505210838544172e-08);         // 1/11!
  p = fma(p, r, 2.755731922398589e-07);         // 1/10!
  p = fma(p, r, 2.7557319223985893e-06);        // 1/9!
  p = fma(p, r, 2.48015873015873e-05);          // 1/8!
  p = fma(p, r, 1.984126984126984e-04);         // 1/7!
  p = fma(p, r, 1.388888888888889e-03);         // 1/6!
  p = fma(p, r, 8.333333333333333e-03);         // 1/5!
  p = fma(p, r, 4.1666666666666664e-02);        // 1/4!
  p = fma(p, r, 1.6666666666666666e-01);        // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const int n = __double2loint(nm);             // low word of the magic sum holds the integer
  const double e = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
  const double d = 1.0 + e;
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  // measured on B200 (tools/pipe_peaks.cu): seed error 9.3e-7, after one Newton step 8.6e-13, after two 1.1e-16
  double err = fma(-d, y, 1.0);
  y = fma(y, err, y);
  err = fma(-d, y, 1.0);
  y = fma(y, err, y);
  return fma(8.0, y, -4.0);
}

// 2^(j/32), j = 0..31, correctly rounded; kernels copy it to shared memory (per-lane indices would serialise in the
// constant cache)
static __constant__ double c_exp2_tbl[32] = {
    1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237,
    1.0905077326652577, 1.1143867425958924, 1.1387886347566916, 1.1637248587775775,
    1.189207115002721, 1.215247359980469, 1.241857812073484, 1.2690509571917332,
    1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832,
    1.4142135623730951, 1.4451808069770467, 1.4768261459394993, 1.5091644275934228,
    1.5422108254079407, 1.5759808451078865, 1.6104903319492543, 1.645755478153965,
    1.681792830507429, 1.718619298122478, 1.7562521603732995, 1.7947090750031072,
    1.8340080864093424, 1.8741676341103, 1.9152065613971474, 1.9571441241754002};

// act_f with a table-driven exp: t = n*ln2 + j*ln2/32 + r, |r| <= ln2/64, exp(t) = 2^n * T[j] * P6(r)
// (degree-6 Taylor: truncation 3.5e-18). 15 fp64-pipe operations for the exponential instead of 22; everything else
// as in act_f. `tbl` is the shared-memory copy of c_exp2_tbl.
__device__ __forceinline__ double act_f_tbl(double x, const double *tbl) {
  double t = -0.5 * x;
  t = fmin(fmax(t, -60.0), 60.0);
  const double magic = 6755399441055744.0;
  const double nm = fma(t, 46.16624130844683, magic);   // 32 / ln 2
  const double nd = nm - magic;
  double r = fma(nd, -0.02166084938653512, t);          // ln2_hi / 32 (exact scaling of fdlibm's split)
  r = fma(nd, -5.9631716539705866e-12, r);              // ln2_lo / 32
  const int ki = __double2loint(nm);
  const double tj = tbl[ki & 31];
  double p = 1.3888888888888889e-03;                    // 1/6!
  p = fma(p, r, 8.333333333333333e-03);
  p = fma(p, r, 4.1666666666666664e-02);
  p = fma(p, r, 1.6666666666666666e-01);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  p *= tj;
  const double e = __hiloint2double(__double2hiint(p) + ((ki >> 5) << 20), __double2loint(p));
  const double d = 1.0 + e;
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double err = fma(-d, y, 1.0);
  y = fma(y, err, y);
  err = fma(-d, y, 1.0);
  y = fma(y, err, y);
  return fma(8.0, y, -4.0);
}
// f'(x) = MAX * (1 - sig^2) / WIDTH with sig = modifiedSigmoid(x/WIDTH) = f(x)/MAX, so it is a function of a = f(x):
__device__ __forceinline__ double act_df_from_a(double a) {
  const double sig = a / 4.0;                        // exact: a = 4 * sig
  const double s2 = __dmul_rn(sig, sig);             // no FMA contraction: the x86 reference rounds sig*sig first
  return __dmul_rn(4.0, __dadd_rn(1.0, -s2)) / 4.0;  // network.h:90
}

// Philox4x32-10 (Salmon et al. 2011); mirrors oracle/mlp_oracle.c:orc_philox4x32_10 bit for bit
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                       uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t) 0xD2511F53u * c0, p1 = (uint64_t) 0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t) (p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t) p1;
    const uint32_t n2 = (uint32_t) (p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t) p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__host__ __device__ __forceinline__ double u53(uint32_t lo, uint32_t hi) {
  const uint64_t v = ((uint64_t) hi << 32) | lo;
  return (double) (v >> 11) * (1.0 / 9007199254740992.0);
}

}  // namespace cab
