/* chessai_b200.h — the C ABI of the B200-native MLP evaluator / trainer for utk003/Chess-AI.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types. The reference has no FFI layer —
 * its boundary is the C++ class API of src/mcts_network/{decider,network}.h — so every entry point below names
 * the reference member it replaces (file:line under /root/reference/src/). The C++ shim classes in
 * chess-ai_b200/host/ (network::Network, network::Optimizer, network::NetworkStorage, decider::Decider) keep the
 * reference signatures and call only this ABI; INTEGRATION.md shows how they slot into src/player and the MCTS loop.
 *
 * Conventions
 *   - every function returns int: 0 = CAB_OK, otherwise a CAB_E* code; the message is in cab_last_error()
 *     (thread-local). No aborts, no exceptions across the ABI (the reference DEBUG_ASSERTs instead).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns CAB_ENODEVICE.
 *   - boards travel as 64 int8 "codes" k in [-9, 9], square i = row*8+col (chess/game.h:121-127), row 0 = white
 *     back rank; the network input is k*0.4 bit-for-bit (chess/piece.cpp:103-105,145-147,179-181,260-262).
 *     k: 1 K, 2 K(moved), 3 Q, 4 R, 5 R(moved), 6 N, 7 B, 8 P, 9 P(moved2x); black negative; 0 empty.
 *   - piece records (encoder input) are 3 bytes per square: type {0 none,1 king,2 queen,3 rook,4 knight,5 bishop,
 *     6 pawn}, colour {0 none,1 white,2 black}, flag {0,1} (King/Rook::_moved, Pawn::_moved2x).
 *   - weights travel as one flat fp64 array in n,i,j order — exactly line 3 of the network_dump text format
 *     (mcts_network/network.cpp:418-422).
 *   - *_host entry points take HOST pointers and include the H2D/D2H copies; *_dev take DEVICE pointers and a
 *     cudaStream_t (passed as void*) and are asynchronous. A null stream means the legacy default stream for the fp64
 *     evaluators (cab_eval_dev, cab_encode_dev); the tensor-core evaluators (cab_eval_bf16_dev, cab_eval_f16_dev,
 *     cab_grad_bf16_dev) and the training entry points then run on the net's OWN non-blocking stream, which the
 *     legacy default stream does not order against: pass a real stream when you time or chain them.
 *   - a cab_net is not thread-safe for mutation (train/dropout/set_weights); eval entry points may be called
 *     concurrently from many host threads on one net (mcts_network/tree.cpp:185 does exactly that). The fp64 and
 *     fused-fp16 evaluators then run concurrently (per-caller scratch); calls of the bf16 wide-net path
 *     (cab_eval_bf16_*, cab_train_batch_bf16_*, cab_grad_bf16_dev) share per-net scratch and TAKE TURNS: they are
 *     serialised on the host by a per-net lock and on the device by an event between consecutive calls.
 */
#ifndef CHESSAI_B200_H_
#define CHESSAI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CAB_API __attribute__((visibility("default")))
#else
#define CAB_API
#endif

#define CAB_OK 0
#define CAB_EINVAL 1    /* bad argument (null pointer, bad dims, code outside [-9,9] is NOT checked on device) */
#define CAB_ENODEVICE 2 /* no CUDA device / driver: the product has no CPU fallback */
#define CAB_ECUDA 3     /* a CUDA runtime call failed; see cab_last_error() */
#define CAB_EIO 4       /* file missing or malformed network_dump / case text */
#define CAB_ENCCL 5     /* NCCL missing or a collective failed */
#define CAB_ESTATE 6    /* call not valid in this state (e.g. dp step before cab_dp_init) */

typedef struct cab_net cab_net; /* opaque: one network resident on one device */

/* ---- library ------------------------------------------------------------------------------------------------ */
CAB_API int cab_version(void);                 /* ABI version, currently 1 */
CAB_API const char *cab_last_error(void);      /* thread-local message of the last failing call */
CAB_API int cab_device_count(int *count);

/* ---- network lifetime (network::Network ctors, clone)   network.cpp:37-79,155-208 ---------------------------- */
/* dims are normalised like Network(std::vector<int>) (network.cpp:39-56): 64 prepended / 1 appended when missing;
 * an empty or <=2-layer result yields the default {64,144,225,100,1}. Weights start at zero (uniformNetwork). */
CAB_API int cab_net_create(const int *dims, int n_dims, int device, cab_net **out);
CAB_API int cab_net_destroy(cab_net *net);
CAB_API int cab_net_clone(const cab_net *net, cab_net **out);               /* Network::clone  network.cpp:195-208 */
CAB_API int cab_net_num_layers(const cab_net *net, int *n_layers);
CAB_API int cab_net_dims(const cab_net *net, int *dims_out);                /* n_layers ints */
CAB_API int cab_net_num_weights(const cab_net *net, long *n_weights);
CAB_API int cab_net_device(const cab_net *net, int *device);

/* ---- weights -------------------------------------------------------------------------------------------------- */
CAB_API int cab_net_set_weights(cab_net *net, const double *w_host, long n);  /* flat n,i,j */
CAB_API int cab_net_get_weights(const cab_net *net, double *w_host, long n);
/* Network::randomNetwork (network.cpp:167-173): U(-1,1). The reference draws from a sequential std::mt19937; here
 * weight w is Philox4x32-10(key = seed, counter = (w, stream)) so the result is order- and device-independent. */
CAB_API int cab_net_randomize(cab_net *net, uint64_t seed, uint64_t stream);
CAB_API int cab_net_zero(cab_net *net);                                        /* Network::uniformNetwork  network.cpp:187-193 */

/* ---- network_dump text format (operator>> / operator<<  network.cpp:386-426, loadFromFile :210-220) ----------- */
/* Reads any dims (the reference's own reader only survives default-shaped files). */
CAB_API int cab_net_load_text(const char *path, int device, cab_net **out);
/* Byte-exact writer: "<L>\n<d0> ... <dL-1>\n" + "%.17e " per weight, no trailing newline (string_util.cpp:39-43). */
CAB_API int cab_net_dump_text(const cab_net *net, const char *path);

/* ---- encoder (Network::loadBoard + Piece::code  network.cpp:119-123) ------------------------------------------ */
CAB_API int cab_encode_host(int device, const uint8_t *records_host, long n_boards, int8_t *codes_host);
/* Device buffers must be 16-byte aligned (any cudaMalloc pointer, any whole-board offset into one). */
CAB_API int cab_encode_dev(const uint8_t *records_dev, long n_boards, int8_t *codes_dev, void *stream);

/* ---- forward (Network::predictPosition  network.cpp:92-117), fp64 reference-precision path --------------------- */
CAB_API int cab_eval_host(cab_net *net, const int8_t *codes_host, long n_boards, double *out_host);
CAB_API int cab_eval_dev(cab_net *net, const int8_t *codes_dev, long n_boards, double *out_dev, void *stream);
/* The leaf-queue shape: n boards already on the device, evaluated as ceil(n/batch) launches of `batch` boards issued
 * round-robin over the caller's streams (cudaStream_t[n_streams]) so consecutive launches overlap. Asynchronous. */
CAB_API int cab_eval_dev_batches(cab_net *net, const int8_t *codes_dev, long n_boards, long batch, double *out_dev,
                                 void *const *streams, int n_streams);
/* encoder fused in front: piece records in, evaluations out */
CAB_API int cab_eval_records_host(cab_net *net, const uint8_t *records_host, long n_boards, double *out_host);
/* forward that also returns every layer's activations (Network::propagate_for_training  network.cpp:139-153):
 * acts_host receives n_boards * sum(dims[1..]) doubles, board-major, layers concatenated in order. */
CAB_API int cab_forward_full_host(cab_net *net, const int8_t *codes_host, long n_boards, double *acts_host);
/* the same forward returning ALSO the pre-activations theta the reference's second overload fills
 * (propagate_for_training(unbounded): unbounded[n][j] = sum, network.cpp:139-153), same shape as acts_host
 * (acts_host may be null). Saturated units keep their true theta (hundreds to thousands with latest.txt). */
CAB_API int cab_forward_thetas_host(cab_net *net, const int8_t *codes_host, long n_boards, double *acts_host,
                                    double *thetas_host);

/* ---- MCTS leaf queue with virtual-loss batching (new; sits in front of Decider::prediction, tree.cpp:295) --------- */
/* Search threads park a playout at an unexpanded node (virtual loss on its path), push the boards they need
 * evaluated and go on with other playouts / trees; one flush evaluates everything queued in launches of `batch`
 * boards over `n_streams` streams. push is thread-safe and lock-free; flush / results are single-threaded. */
typedef struct cab_leafq cab_leafq;
CAB_API int cab_leafq_create(cab_net *net, long capacity, long batch, int n_streams, cab_leafq **out);
CAB_API int cab_leafq_destroy(cab_leafq *q);
CAB_API int cab_leafq_push(cab_leafq *q, const int8_t *codes, long n_boards, long *first_slot);
CAB_API int cab_leafq_pending(cab_leafq *q, long *n_boards);
/* 0 = fp64 reference precision (default), 1 = the fused fp16 tcgen05 kernel (reduced precision, cab_eval_f16_*) */
CAB_API int cab_leafq_set_precision(cab_leafq *q, int precision);
CAB_API int cab_leafq_flush(cab_leafq *q, long *n_evaluated);
/* The same in two halves: flush_async queues H2D + launches + D2H and returns; wait blocks until the results are valid and
 * empties the queue. No push in between. A search thread with two queues overlaps its tree walk with the GPU. */
CAB_API int cab_leafq_flush_async(cab_leafq *q, long *n_evaluated);
CAB_API int cab_leafq_wait(cab_leafq *q);
CAB_API int cab_leafq_results(cab_leafq *q, const double **results_by_slot);
CAB_API int cab_leafq_stats(cab_leafq *q, long *launches, long *flushes, long *evaluated);
/* K7 — priors of expanded nodes (MCTS::expand_node, tree.cpp:299-306 + Node::expand :92): node s owns
 * logits[seg_off[s] .. seg_off[s+1]); prior_i = exp(cm_s * logit_i) / sum over the node. */
/* Decider::prediction (decider.cpp:31-61) for a position on the network's code board (side_to_move = +1 white, -1
 * black): *eval_out = cm * net(position); per legal move, in the reference's std::map order ((from, to) ascending, the
 * four promotions of a pawn move sharing one key whose value is the last one assigned), moves_out[2i] = from square,
 * moves_out[2i+1] = to square, logits_out[i] = its logit. The position and all children that need the network are
 * evaluated in ONE batched launch. network_priors = 0: the shipped branch (cm * 20 unless the opponent has no reply,
 * then cm * net(child)); 1: cm * net(child) for every move. *n_moves = number of keys (may exceed cap: truncated). */
CAB_API int cab_prediction_host(cab_net *net, const int8_t *codes64, int side_to_move, int network_priors, double *eval_out,
                        int *moves_out, double *logits_out, int cap, int *n_moves);
/* The move the network ranks first (first child in std::map order with the greatest net(child): the weight of a move
 * is exp(cm * logit) = exp(net(child)), decider.cpp:52-56 + tree.cpp:302), with the reduced-precision evaluators made
 * safe for ranking: precision 0 = fp64 (the reference answer); 1 = fused fp16 tcgen05; 2 = bf16 tcgen05 (wide nets).
 * For 1 / 2 every child within `tol` of the reduced-precision maximum is re-evaluated in fp64 and the maximum taken
 * over the refined values, so the move is IDENTICAL to the fp64 choice whenever the reduced path errs by <= tol / 2
 * per board (tol <= 0: 0.1 for fp16, 0.3 for bf16 = twice the stated tolerances). *n_refined = boards redone in fp64. */
CAB_API int cab_best_move_host(cab_net *net, const int8_t *codes64, int side_to_move, int precision, double tol, int *from_out,
                               int *to_out, double *value_out, int *n_moves, int *n_refined);
CAB_API int cab_prior_softmax_host(int device, const double *logits, const int *seg_off, int n_seg, const double *color_mult,
                                   double *priors_out);

/* ---- MCTS self-play resident on the GPU (new; tree::MCTS + Decider::prediction + Game::tryMove as kernels) ----------- */
/* BASELINE.json configs[2]. The trees (one per game and root), the playouts, the node expansion — legal moves in the
 * reference's order (chess/game.cpp:228-263), child boards written straight into the evaluator's input buffer, the
 * "opponent has no reply" test (mcts_network/decider.cpp:47-57) — the priors (tree.cpp:299-306), the virtual losses, the
 * move choice (tree.cpp:122-142,195) and the game rules (game.cpp:847-890) all run on the device; per round the host
 * launches three kernels and reads one counter. Search semantics are tree.cpp's: with in_flight = 1, roots = 1 and
 * noise = 0 a search equals tree::MCTS::mcts() (visit counts exactly; tests/test_mcts_gpu.py). */
typedef struct cab_mcts cab_mcts;
typedef struct cab_mcts_params {
  int games;              /* concurrent games on this device */
  int roots;              /* independent trees per move, merged like Node::combineNodes (DEFAULT_NUM_THREADS, tree.cpp:147) */
  int playouts_per_move;  /* split over the roots (NUM_SIMULATIONS_PER_THREAD = playouts_per_move / roots) */
  int in_flight;          /* playouts in flight per tree under virtual loss; 1 = the reference's sequential search */
  int max_moves;          /* a game stops after this many moves (or at mate / stalemate / 50 plies without capture) */
  int network_priors;     /* 0: decider.cpp:52-57 as shipped (constant logit 20 unless the opponent has no reply); 1: every child evaluated */
  int noise;              /* 1: Dirichlet(0.2) noise mixed 25 % into the root priors (tree.cpp:318-347), Philox-driven */
  int precision;          /* 0: fp64 evaluator (the reference's search); 1: the fused fp16 tcgen05 evaluator (cab_eval_f16_*: faster, NOT a parity path) */
  double save_ratio;      /* probability of keeping the board after a move as a training case (initialization.cpp:221-227) */
  uint64_t seed;
} cab_mcts_params;
typedef struct cab_mcts_stats {
  long playouts, expansions, boards_evaluated, moves_played, games_finished, searches_done, rounds;
  long arena_full, eval_full, cases_kept;
  double seconds;            /* wall time inside cab_mcts_run */
  double evaluator_seconds;  /* of which the network evaluator was running (CUDA events around its launches) */
} cab_mcts_stats;
CAB_API int cab_mcts_create(cab_net *net, const cab_mcts_params *params, cab_mcts **out);
CAB_API int cab_mcts_destroy(cab_mcts *q);
/* start positions (default: the start position): codes [games][64], side [games] (+1 white / -1 black to move),
 * no_capture [games] or null. Restarts every game. */
CAB_API int cab_mcts_set_positions(cab_mcts *q, const int8_t *codes, const int8_t *side, const int *no_capture);
/* search_only = 0: play every game to its end; 1: search the current move of every game and stop (nothing is played).
 * max_rounds > 0 bounds the call (call again to continue). */
CAB_API int cab_mcts_run(cab_mcts *q, int search_only, long max_rounds, cab_mcts_stats *stats);
/* searched root of one game: children in std::map<Move> order with from / to square, prior, visit count, value sum;
 * root = -1 merges the roots (Node::combineNodes), root >= 0 reads one tree. */
CAB_API int cab_mcts_root_children(cab_mcts *q, int game, int root, int cap, int *from, int *to, double *prior, int *visits,
                                   double *value_sum, int *n_children, int *root_visits, double *root_value_sum);
CAB_API int cab_mcts_game_state(cab_mcts *q, int game, int8_t *codes64, int *side, int *over, int *result, int *moves_played);
/* training cases of the finished games: target = 10 * (4 * clamp(2v - 1) + result) / 5 (make_cases.cpp:36-42) */
CAB_API int cab_mcts_cases(cab_mcts *q, int8_t *codes_out, double *targets_out, long cap, long *n);
/* the kernels' own move generator on n positions (parity tests against the recorded reference games):
 * moves_out [n][cap][3] = (from, to, promo), reference order (game.cpp:228-263, promotions Q R N B unless queen_only);
 * no_reply_out (optional, needs queen_only) [n][cap]: 1 where the move leaves the opponent no legal move */
CAB_API int cab_movegen_host(int device, const int8_t *codes, const int8_t *side, int n, int queen_only, int cap,
                             int *n_moves_out, uint8_t *moves_out, uint8_t *no_reply_out);

/* ---- bf16 tensor-core forward (tcgen05 + TMEM + TMA) for WIDE networks ------------------------------------------ */
/* Same forward (network.cpp:92-117) with bf16 operands, fp32 accumulation in tensor memory and f() in fp32. Only for
 * networks whose hidden widths are multiples of 256 (e.g. Network({2048,2048,2048,2048})); CAB_EINVAL otherwise.
 * Stated tolerance vs the fp64 path: |dy| <= 0.15 on the (-4, 4) output scale for weights U(-1,1)/sqrt(fan_in),
 * identical sign of the evaluation away from |y| < 0.15 (tests/test_tc_gpu.py). */
CAB_API int cab_eval_bf16_dev(cab_net *net, const int8_t *codes_dev, long n_boards, double *out_dev, void *stream);
CAB_API int cab_eval_bf16_host(cab_net *net, const int8_t *codes_host, long n_boards, double *out_host);
/* Optional REDUCED-PRECISION fast path for nets shaped {64, N1, N2, N3, 1}, N1..N3 <= 256 (the default network):
 * Network::predictPosition as ONE tcgen05 kernel, fp16 operands, fp32 accumulation, weights resident in shared
 * memory, activations kept in tensor memory between the layers. NOT a parity path. Stated tolerance vs the fp64 path:
 * |dy| <= 0.05 on the (-4, 4) output scale for weights U(-1,1)/sqrt(fan_in) and identical sign wherever |y| >= 0.05;
 * for the shipped, saturated latest.txt: the same sign on >= 99.5 % of synthetic boards (tests/test_tc_gpu.py). */
CAB_API int cab_eval_f16_dev(cab_net *net, const int8_t *codes_dev, long n_boards, double *out_dev, void *stream);
CAB_API int cab_eval_f16_host(cab_net *net, const int8_t *codes_host, long n_boards, double *out_host);
/* n boards as launches of `batch`, round-robin over the caller's streams (like cab_eval_dev_batches) */
CAB_API int cab_eval_f16_dev_batches(cab_net *net, const int8_t *codes_dev, long n_boards, long batch, double *out_dev,
                             void *const *streams, int n_streams);

/* ---- training (network::Optimizer  network.cpp:223-292, train loop main/network/train.cpp:39-68) ---------------- */
/* Optimizer::optimize on ONE case: forward, backprop with in-place update (step lambda), re-forward, accept
 * (lambda *= rate while lambda < max_const) or reject (lambda /= rate and roll the update back). */
CAB_API int cab_train_case(cab_net *net, const int8_t *codes64_host, double target, double *lambda, double max_const,
                   double rate, int *accepted, double *err, double *new_err);
/* train_network's loop body over n cases in the given order, on the device in one launch sequence: optimize, then
 * if lambda leaves [1e-4, 10]: lambda = 2.0 and dropout(rate 0.01) with Philox(seed, offset = reset index).
 * accepted_host (optional) receives n flags; resets (optional) the number of lambda resets. */
CAB_API int cab_train_sequence(cab_net *net, const int8_t *codes_host, const double *targets_host, long n_cases,
                       double *lambda, uint64_t dropout_seed, uint8_t *accepted_host, int *resets);
/* Minibatch extension of optimize (no reference counterpart for B > 1; reduces to it at B == 1): dW is the batch
 * MEAN of the per-sample updates, E and E' are batch means, one accept/reject per step. With cab_dp_init the sums
 * are all-reduced over ranks first, so every rank takes the identical decision and update. */
CAB_API int cab_train_batch_host(cab_net *net, const int8_t *codes_host, const double *targets_host, long n_cases,
                         double *lambda, double max_const, double rate, int *accepted, double *err, double *new_err);
CAB_API int cab_train_batch_dev(cab_net *net, const int8_t *codes_dev, const double *targets_dev, long n_cases,
                        double *lambda, double max_const, double rate, int *accepted, double *err, double *new_err,
                        void *stream);
/* train_network's loop (train.cpp:39-68) over n_cases RESIDENT cases as consecutive minibatches of `batch` (the last may be
 * shorter) with NO host round trip inside: lambda, the accept / reject decision, the rollback and (reset_rule != 0) the
 * lambda-reset + dropout rule of train.cpp:54-57 live on the device; every step is enqueued behind the previous one and
 * the call waits once, at the end. Identical arithmetic to calling cab_train_batch_dev per minibatch (tests).
 * e_trace_host (optional): 2 doubles per step (E, E'). */
typedef struct cab_train_stats {
  long steps, accepted;
  int resets, last_accepted;
  double first_err, last_err, last_new_err;
} cab_train_stats;
CAB_API int cab_train_epoch_dev(cab_net *net, const int8_t *codes_dev, const double *targets_dev, long n_cases, long batch,
                                double *lambda, double max_const, double rate, int reset_rule, uint64_t dropout_seed,
                                cab_train_stats *stats, double *e_trace_host, void *stream);
/* the same with the cases in HOST memory: one upload, then the resident epoch (what the trainer command line calls per
 * save interval) */
CAB_API int cab_train_epoch_host(cab_net *net, const int8_t *codes_host, const double *targets_host, long n_cases, long batch,
                                 double *lambda, double max_const, double rate, int reset_rule, uint64_t dropout_seed,
                                 cab_train_stats *stats, double *e_trace_host);
CAB_API int cab_train_epoch_bf16_dev(cab_net *net, const int8_t *codes_dev, const double *targets_dev, long n_cases, long batch,
                                     double *lambda, double max_const, double rate, int reset_rule, uint64_t dropout_seed,
                                     cab_train_stats *stats, double *e_trace_host, void *stream);
/* The same minibatch step on the bf16 tcgen05 path, for WIDE nets (hidden widths multiples of 256; BASELINE.json
 * configs[4]): forward with saved activations, dX and dW as tensor-core GEMMs (fp32 accumulate, psi kept in bf16),
 * gradient in fp64, update / re-forward / accept-or-roll-back exactly as above on the fp64 master weights.
 * Stated tolerance vs the fp64 rule: relative L2 error of each layer's dW <= 3e-2 (tests/test_tc_gpu.py). */
CAB_API int cab_train_batch_bf16_host(cab_net *net, const int8_t *codes_host, const double *targets_host, long n_cases,
                              double *lambda, double max_const, double rate, int *accepted, double *err, double *new_err);
CAB_API int cab_train_batch_bf16_dev(cab_net *net, const int8_t *codes_dev, const double *targets_dev, long n_cases,
                             double *lambda, double max_const, double rate, int *accepted, double *err, double *new_err,
                             void *stream);
/* forward + backward only: sum_b dW into cab_grad_buffer()[0..nw), then sum_b E, unused, count */
CAB_API int cab_grad_bf16_dev(cab_net *net, const int8_t *codes_dev, const double *targets_dev, long n_cases, void *stream);
/* Optimizer::dropout ("weak dilution", network.cpp:281-292): each weight w.p. rate -> U(-1,1), Philox(seed, offset). */
CAB_API int cab_dropout(cab_net *net, double rate, uint64_t seed, uint64_t offset, long *hits);

/* ---- data-parallel training: one process per GPU, NCCL all-reduce of (sum dW, sum E, sum E') ------------------- */
CAB_API int cab_dp_unique_id(void *id128_out);                                   /* rank 0: ncclGetUniqueId (128 bytes) */
CAB_API int cab_dp_init(cab_net *net, const void *id128, int rank, int world);   /* ncclCommInitRank on the net's device */
CAB_API int cab_dp_shutdown(cab_net *net);
/* The same data-parallel step WITHOUT a library collective on its critical path: every rank's gradient buffer is mapped
 * into every other rank's address space (CUDA IPC; the ranks' devices must have peer access: NVLink / NVSwitch) and ONE
 * kernel per step waits for the ranks' flags, sums the ranks' gradients in rank order (identical bits everywhere) and
 * applies the update; E' crosses as 8-byte peer stores inside the decision kernel. One process per GPU:
 *   every rank: cab_dp_peer_export -> exchange the 64-byte handles (any transport) -> cab_dp_peer_attach(all, rank, world)
 * then cab_train_batch_* / cab_train_epoch_* as usual; cab_dp_shutdown detaches (after a barrier of the caller's).
 * A peer that never signals makes the step return CAB_ESTATE after ~2 s instead of hanging. 2 <= world <= 16. */
CAB_API int cab_dp_peer_export(cab_net *net, void *handle64_out);
CAB_API int cab_dp_peer_attach(cab_net *net, const void *handles64, int rank, int world);
/* raw device pointers of the gradient / scalar buffers, for callers that run their own collective
 * (e.g. torch.distributed): grad has n_weights doubles followed by 2 doubles (sum E, sum E') and 1 count. */
CAB_API int cab_grad_buffer(cab_net *net, double **grad_dev, long *n_doubles);

/* ---- measurement helpers ---------------------------------------------------------------------------------------- */
/* kind 0: fp64 tensor pipe (DMMA m8n8k4), 1: fp64 FMA pipe, 2: fp32 FMA pipe. Best of 10, CUDA events. */
CAB_API int cab_measure_pipe_peak(int device, int kind, double *tflops);
/* number of kernels this library has launched on behalf of `net` since creation (bench.py's gpu_launches) */
CAB_API int cab_net_launch_count(const cab_net *net, long *launches);

#ifdef __cplusplus
}
#endif
#endif /* CHESSAI_B200_H_ */
