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

// kDot: the chunk is a one-output layer evaluated as a dot product on the previous layer's activations still in
// registers (forward kernel); kPreDot: last chunk of the layer in front of it (its activations are not stored)
enum ChunkFlags : uint8_t { kFirst = 1, kLast = 2, kFinal = 4, kDot = 8, kPreDot = 16 };

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
  uint8_t *d_blob = nullptr;  // [ChunkDesc x kSmemChunks][exp table]: what the forward kernel copies to shared memory at start
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
  // stale-copy flags: set by whoever changes d_w, tested and cleared only under `mu` (evaluation) / `tc_mu` (bf16) and
  // only after the stream that rebuilt the copy has drained, so a launch on another stream never overtakes the rebuild
  std::atomic<bool> pack_dirty{true};    // packed streams stale w.r.t. d_w
  bool fp64_ok = true;       // the fp64 stream kernels fit this network's widest layer in shared memory
  std::atomic<bool> tc_dirty{true};      // bf16 copies of the tcgen05 path stale w.r.t. d_w
  std::atomic<bool> f16_dirty{true};     // fp16 shared-memory image of the fused tcgen05 path stale w.r.t. d_w
  std::atomic<void *> f16{nullptr};      // F16State of cab_tc.cu, published only when fully constructed
  void *tc = nullptr;        // cab::TcState of the bf16 tcgen05 path (cab_tc.cu), created on first use
  std::mutex tc_mu;          // the bf16 path keeps per-net scratch (activations, staging): its calls take turns
  cudaEvent_t tc_done = nullptr;  // end of the previous bf16 call: the next one, on whatever stream, waits for it
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
  void *d_ctl = nullptr;       // cab::TrainCtl: the minibatch trainer's lambda / decision / counters, resident on the device
  cudaStream_t train_stream = nullptr;
  cudaStream_t aux_stream[4] = {};  // helper streams for fanned-out passes (cab_train.cu:fan_out)
  cudaEvent_t aux_event[5] = {};
  // data parallel
  void *nccl_comm = nullptr;
  int dp_rank = 0, dp_world = 1;
  // data parallel over NVLink peer memory (cab_dp_peer_*): every rank's exchange block (two gradient buffers, E' slots,
  // epoch flags) is mapped into every other rank's address space through CUDA IPC
  void *peer_block = nullptr;              // this rank's block
  void *peer_map[16] = {};                 // rank r's block in THIS address space (peer_map[dp_rank] == peer_block)
  void **d_peer_table = nullptr;           // the same table on the device
  double *d_grad_own = nullptr;            // the private gradient buffer d_grad points to outside peer mode
  unsigned long long peer_epoch = 0;       // steps taken in peer mode (the flags' values)
  int peer_world = 0;                      // > 1: the trainer reduces over peer memory instead of NCCL
  // eval contexts
  std::mutex mu;
  std::vector<cab::EvalCtx *> ctxs;
  std::atomic<long> launches{0};
};

namespace cab {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

void dp_peer_detach(cab_net *net);   // cab_train.cu

// cab_api.cu, used by the other translation units
int ensure_packed(cab_net *net, cudaStream_t stream, bool drain);
int launch_forward(cab_net *net, const int8_t *d_codes, long n_boards, double *d_out, double *d_save, long save_stride,
                   cudaStream_t stream, double *d_theta = nullptr, bool alone = false);
struct Geometry { int n_cons, n_stages, ctas_per_sm; };
Geometry choose_geometry(int a_units, int stage_bytes, long n_boards, int sm_count);
size_t forward_smem_bytes(int a_units, int n_cons, int n_stages, int stage_bytes);

#define CAB_CUDA(call)                                                         \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) return cab::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)

// cudaMemset of device memory on the legacy stream returns before the fill has run, and the library's own streams are
// created cudaStreamNonBlocking: they do NOT wait for the legacy stream. A buffer zeroed at allocation and written right
// after on such a stream (the trainer's lambda upload into d_scalars) could be zeroed AFTER the write (seen once in ~10
// runs of lib/train on a slow box: lambda read as 0 -> reset rule -> dropout). Every set-up fill therefore completes here.
inline cudaError_t memset_now(void *p, int value, size_t bytes) {
  cudaError_t e = cudaMemset(p, value, bytes);
  return e != cudaSuccess ? e : cudaStreamSynchronize(cudaStreamLegacy);
}

// Likewise a synchronous cudaMemcpy from PAGEABLE host memory returns once the bytes sit in the driver's staging buffer;
// the DMA to the device runs on the legacy stream. Set-up uploads (weights, plans, tables) complete here before any
// non-blocking stream may read them.
inline cudaError_t upload_now(void *dst, const void *src, size_t bytes) {
  cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
  return e != cudaSuccess ? e : cudaStreamSynchronize(cudaStreamLegacy);
}

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
static __constant__ double c_exp2_tbl[256] = {
    1.0, 1.0027112750502025, 1.0054299011128027, 1.0081558981184175,
    1.0108892860517005, 1.0136300849514894, 1.016378314910953, 1.019133996077738,
    1.0218971486541166, 1.0246677928971357, 1.0274459491187637, 1.030231637686041,
    1.0330248790212284, 1.0358256936019572, 1.0386341019613787, 1.041450124688316,
    1.0442737824274138, 1.0471050958792898, 1.0499440858006872, 1.0527907730046264,
    1.0556451783605572, 1.0585073227945128, 1.061377227289262, 1.0642549128844645,
    1.0671404006768237, 1.0700337118202419, 1.0729348675259756, 1.075843889062791,
    1.0787607977571199, 1.0816856149932152, 1.0846183622133092, 1.0875590609177697,
    1.0905077326652577, 1.0934643990728858, 1.0964290818163769, 1.099401802630222,
    1.102382583307841, 1.1053714457017412, 1.1083684117236787, 1.1113735033448175,
    1.1143867425958924, 1.1174081515673693, 1.1204377524096067, 1.12347556733302,
    1.1265216186082418, 1.129575928566288, 1.1326385195987192, 1.1357094141578055,
    1.1387886347566916, 1.1418762039695616, 1.1449721444318042, 1.148076478840179,
    1.1511892299529827, 1.154310420590216, 1.1574400736337511, 1.1605782120274988,
    1.1637248587775775, 1.1668800369524817, 1.1700437696832502, 1.1732160801636373,
    1.1763969916502812, 1.1795865274628758, 1.182784710984341, 1.1859915656609938,
    1.189207115002721, 1.1924313825831512, 1.1956643920398273, 1.1989061670743806,
    1.202156731452703, 1.2054161090051239, 1.2086843236265816, 1.2119613992768012,
    1.215247359980469, 1.2185422298274085, 1.2218460329727576, 1.2251587936371455,
    1.22848053610687, 1.2318112847340759, 1.2351510639369334, 1.2384998981998165,
    1.241857812073484, 1.245224830175258, 1.2486009771892048, 1.2519862778663162,
    1.255380757024691, 1.2587844395497165, 1.2621973503942507, 1.2656195145788063,
    1.2690509571917332, 1.2724917033894028, 1.275941778396392, 1.2794012075056693,
    1.2828700160787783, 1.2863482295460256, 1.2898358734066657, 1.2933329732290895,
    1.2968395546510096, 1.3003556433796506, 1.3038812651919358, 1.3074164459346773,
    1.3109612115247644, 1.3145155879493546, 1.318079601266064, 1.3216532776031575,
    1.3252366431597413, 1.3288297242059544, 1.3324325470831615, 1.3360451382041458,
    1.339667524053303, 1.3432997311868353, 1.3469417862329458, 1.3505937158920345,
    1.3542555469368927, 1.3579273062129011, 1.3616090206382248, 1.365300717204012,
    1.3690024229745905, 1.3727141650876684, 1.3764359707545302, 1.380167867260238,
    1.383909881963832, 1.387662042298529, 1.3914243757719262, 1.3951969099662003,
    1.3989796725383112, 1.4027726912202048, 1.4065759938190154, 1.4103896082172707,
    1.4142135623730951, 1.4180478843204152, 1.4218926021691656, 1.4257477441054942,
    1.42961333839197, 1.433489413367789, 1.4373759974489824, 1.4412731191286257,
    1.4451808069770467, 1.449099089642035, 1.4530279958490526, 1.4569675544014438,
    1.460917794180647, 1.4648787441464057, 1.4688504333369818, 1.4728328908693675,
    1.4768261459394993, 1.4808302278224719, 1.4848451658727524, 1.488870989524397,
    1.4929077282912648, 1.4969554117672355, 1.5010140696264256, 1.5050837316234065,
    1.5091644275934228, 1.5132561874526098, 1.5173590411982147, 1.5214730189088146,
    1.5255981507445384, 1.529734466947287, 1.533881997840956, 1.5380407738316568,
    1.5422108254079407, 1.5463921831410214, 1.550584877685, 1.5547889397770887,
    1.559004400237837, 1.5632312899713576, 1.567469639965553, 1.5717194812923414,
    1.5759808451078865, 1.5802537626528246, 1.5845382652524937, 1.588834384317164,
    1.593142151342267, 1.597461597908627, 1.6017927556826934, 1.606135656416771,
    1.6104903319492543, 1.6148568142048607, 1.6192351351948637, 1.6236253270173289,
    1.6280274218573478, 1.632441451987275, 1.6368674497669644, 1.6413054476440063,
    1.645755478153965, 1.6502175739206177, 1.6546917676561943, 1.6591780921616162,
    1.6636765803267364, 1.6681872651305825, 1.6727101796415966, 1.6772453570178785,
    1.681792830507429, 1.6863526334483934, 1.6909247992693053, 1.6955093614893326,
    1.7001063537185235, 1.7047158096580513, 1.709337763100463, 1.713972247929926,
    1.718619298122478, 1.723278947746274, 1.7279512309618377, 1.732636182022311,
    1.7373338352737062, 1.7420442251551564, 1.746767386199169, 1.7515033530318782,
    1.7562521603732995, 1.761013843037584, 1.7657884359332727, 1.7705759740635547,
    1.7753764925265212, 1.7801900265154245, 1.785016611318935, 1.789856282321401,
    1.7947090750031072, 1.7995750249405351, 1.804454167806624, 1.809346539371032,
    1.8142521755003989, 1.8191711121586085, 1.8241033854070534, 1.8290490314048973,
    1.8340080864093424, 1.8389805867758937, 1.843966568958626, 1.8489660695104508,
    1.8539791250833855, 1.8590057724288205, 1.864046048397789, 1.8690999899412386,
    1.8741676341103, 1.8792490180565602, 1.8843441790323345, 1.8894531543909392,
    1.8945759815869656, 1.8997126981765553, 1.9048633418176741, 1.9100279502703899,
    1.9152065613971474, 1.9203992131630474, 1.925605943636125, 1.930826790987627,
    1.9360617934922943, 1.9413109895286405, 1.9465744175792332, 1.9518521162309783,
    1.9571441241754002, 1.9624504802089273, 1.9677712232331759, 1.9731063922552343,
    1.978456026387951, 1.9838201648502194, 1.9891988469672663, 1.9945921121709402};
constexpr int kExpTblSize = 256;

// act_f with a table-driven exp, arranged for the FEWEST fp64-pipe operations (the pipe is shared with the DMMAs, so
// every activation operation is a DMMA slot lost):
//   t = -x/2 = (256 k + j) ln2/256 + r,  |r| <= ln2/512;  exp(t) = 2^k T[j] exp(r),  T[j] = 2^(j/256)
//   * the saturation clamp |x| <= 120 (exp(+-60) already gives f = -+4 exactly like the reference's inf / 0 path) is
//     done on the INTEGER pipe, on the high word of x (also catches inf and NaN)
//   * the kernel works with r' = -2 r = x + n (ln2/128), so t itself is never formed
//   * exp(r) = p = 1 + r' q(r'), q of degree 3 (truncation r^5/120 < 3.8e-17)
//   * d8 = (1 + exp(t)) / 8 = 1/8 + T8 p with T8 = 2^(k-3) T[j] (exponent arithmetic on the integer pipe): ONE fma,
//     the product exact inside it
//   * 8 / (1 + exp(t)) = 1 / d8 by MUFU.RCP64H (relative error 9.3e-7) and ONE cubic step y (1 + e + e^2), e = 1 - d8 y
//     (residual e^3 < 1e-18);  f = 8 / (1 + exp(t)) - 4
// 13 fp64-pipe operations (round 1: 16; the libm expression: ~51 lane-clocks). An fp64 operation holds the issue port
// for two cycles and the integer / MUFU / LDS instructions do not hide under it (tools/act_bench.cu: 40 issue cycles per
// warp activation against a 26-cycle pipe floor), so those are kept few too.
// `tbl` is the shared-memory table T[j] / 8 = 2^(j/256 - 3), j = 0..255 (c_exp2_tbl scaled exactly).
__device__ __forceinline__ double act_f_tbl(double x, const double *tbl) {
  const int hx = __double2hiint(x);
  const bool big = (unsigned) (hx & 0x7fffffff) >= 0x405e0000u;           // |x| >= 120.0 (0x405E000000000000), inf, NaN
  // (the low word stays: +-[120, 120 + 2^-14) saturates exactly like +-120)
  const double xc = __hiloint2double(big ? ((hx & (int) 0x80000000) | 0x405e0000) : hx, __double2loint(x));
  const double magic = 6755399441055744.0;
  const double nm = fma(xc, -184.66496523378732, magic);   // n = rint(-x/2 * 256/ln2)
  const double nd = nm - magic;
  double r = fma(nd, 0.00541521234663378, xc);              // r' = x + n * ln2_hi/128   (fdlibm's split of ln 2, scaled exactly)
  r = fma(nd, 1.4907929134926466e-12, r);                   //        + n * ln2_lo/128
  const int ki = __double2loint(nm);
  const double tj = tbl[ki & 255];                          // T[j] / 8
  const double t8 = __hiloint2double(__double2hiint(tj) + ((ki & (int) 0xffffff00) << 12), __double2loint(tj));   // * 2^k
  double q = fma(r, 2.6041666666666665e-03, -2.0833333333333332e-02);   // 1/384, -1/48
  q = fma(r, q, 0.125);
  q = fma(r, q, -0.5);
  const double pe = fma(r, q, 1.0);
  const double d8 = fma(t8, pe, 0.125);
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d8));
  const double e = fma(-d8, y, 1.0);
  const double e2 = fma(e, e, e);
  return fma(y, e2, y - 4.0);
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
