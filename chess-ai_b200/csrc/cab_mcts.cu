// cab_mcts.cu — MCTS self-play with virtual-loss leaf batching, RESIDENT ON THE GPU (BASELINE.json configs[2]).
//
// Round 1 walked the trees on the host (host/selfplay.cpp) and was bound by the host cores: one B200 evaluated 11 % of
// what it can, eight B200s behind 32 cores scaled 0.35. Here the whole search lives in HBM and runs in kernels:
//   trees        one node arena per tree (game x root), 40-byte nodes, children of a node contiguous
//   walk kernel  ONE WARP PER TREE. It resumes the tree's parked playouts (virtual loss off, priors formed from the
//                evaluations of the previous round: K7, tree.cpp:299-306), then launches new ones: PUCT selection with
//                the children spread over the lanes (tree.cpp:276-283), the move played on the warp's board in shared
//                memory; at an unexpanded node the playout puts a virtual loss on its path, asks for the expansion and
//                parks. Playouts of one tree are walked one after the other (each sees the virtual losses of the
//                ones before it, and the search stays deterministic); the trees run side by side.
//   expand kernel ONE WARP PER REQUEST: legal moves in the reference's order with the from-squares spread over the
//                lanes (chess_rules.cuh), child nodes allocated, the parent and the child boards written straight into
//                the evaluator's input buffer, "opponent has no reply" (decider.cpp:47-57) decided per child.
//   evaluator    mlp_forward_kernel on the boards queued by the round (cab_api.cu launch_forward), results by slot
//   game kernel  per game, once all its trees have spent their playouts: roots merged like Node::combineNodes
//                (tree.cpp:122-142), the move with the best PUCT score played (:195), mate / stalemate / 50-ply rule
//                (game.cpp:847-890), training case kept with probability save_ratio, new roots.
// The host only launches rounds and reads one counter per round; games of different GPUs never talk to each other.
//
// Search semantics are the reference's (tree.cpp, restated; cited per item) exactly as in host/selfplay.cpp and
// integration/batched_mcts.cpp: with one playout in flight and no noise a search reproduces tree::MCTS::mcts() —
// visit counts, value sums and priors (tests/test_mcts_gpu.py against tests/golden/mcts_visits.npz, recorded from the
// reference's own code). Bit-exactness of the PUCT comparisons: the score is computed in double in the reference's
// operation order without FMA contraction; log(...) + 1.25 and the leaf sigmoid come from tables the HOST fills with
// its libm (the same values the reference computes), sqrt and the divisions are IEEE on both sides.
#include <chrono>
#include <cmath>
#include <cstring>
#include <vector>

#include "cab_internal.cuh"
#include "chess_rules.cuh"

namespace cab {
namespace mcts {

constexpr int kDepth = 8;             // tree.cpp:145 SIMULATION_SEARCH_DEPTH
constexpr int kMaxChildren = 256;     // fastchess kMaxMoves
constexpr int kWarpsPerCta = 4;
constexpr long kEvalPiece = 8192;     // boards per evaluator launch when a round's boards are spread over several streams
constexpr int kBoardStride = 68;      // lane-private child boards: 17 words apart, so one square of all lanes hits 32 banks

enum NodeFlags : uint8_t { kExpanded = 1, kPending = 2, kEval = 4 };

struct __align__(8) Node {            // 40 bytes
  double prior, value_sum;
  int visits, vloss;                  // visits includes the virtual losses of parked playouts
  int first_child;
  uint16_t n_children;
  int8_t color;                       // colour to move at this node
  uint8_t flags;
  int8_t state;                       // 0 not examined, 1 the side to move has a legal move, 2 game over here
  int8_t result;                      // winner when state == 2 (0 draw)
  uint8_t from, to, promo;            // the move that leads here (promo: 3 = queen, the std::map<Move> key's promotion)
  uint8_t pad[3];
};
static_assert(sizeof(Node) == 40, "Node must stay 40 bytes");

struct Pos {                          // fastchess::Position
  int8_t sq[64];
  int8_t side, over, result, pad;
  int no_capture;
};
static_assert(sizeof(Pos) == 72, "Pos must stay 72 bytes");

struct Playout {                      // a parked playout
  Pos pos;
  int node, depth, path_len, slot;    // slot: first evaluation slot of the expansion it requested, -1: only waiting
  int path[kDepth + 2];
};

struct TreeState { int n_nodes, launched, done, n_parked; };
struct GameState {
  Pos pos;
  int moves_played, finished, search_done, pad;
};
struct KeptCase {                     // a board kept for training, waiting for its game's result
  int8_t codes[64];
  double bounded;                     // clamp(2 v - 1) of the searched root (initialization.cpp:221-227)
  int game, pad;
};

struct Counters {
  unsigned long long playouts, expansions, boards, moves, games_finished, searches_done, arena_full, eval_full, cases;
};

struct Dev {
  Node *nodes; long node_cap;
  TreeState *trees;
  Playout *parked;
  GameState *games;
  int n_games, roots, quota, in_flight, max_moves, network_priors, noise, stop_after_search, max_launch;
  int8_t *eval_codes[2]; double *eval_out[2]; int *eval_count[2]; int eval_cap;
  int *req; int *req_count;             // expansion requests of the round: index of the parked playout that asked
  const double *pf_tbl; int pf_n;       // log((N + 19652 + 1) / 19652) + 1.25 by parent visit count N (host libm)
  const double *sig_tbl;                // 1 / (1 + exp(-m / 27.18)) by material m + 1024 (host libm)
  double exp20;                         // exp(20.0) (host libm): the constant-logit weight of decider.cpp:56
  Counters *cnt;
  KeptCase *cases; int cases_cap; double save_ratio;
  unsigned long long seed;
};

// ---- per-warp shared memory ---------------------------------------------------------------------------------------
struct WarpShared {
  Pos pos;                                      // the playout being walked
  int path[kDepth + 2];
  uint16_t moves[kMaxChildren];                 // from | to << 8 of the expansion under construction
  uint32_t eval_mask[kMaxChildren / 32];        // children whose board goes to the network
  int8_t base[64];                              // the parent with every moved2x flag cleared: what each child starts from
  int8_t child[32 * kBoardStride];              // one scratch board per lane
  double weight[kMaxChildren];                  // exp(cm * logit) per child while the priors are formed
};

__device__ __forceinline__ int lane_id() { return (int) (threadIdx.x & 31); }

// Philox-based uniforms for the root noise and the case sampling
__device__ __forceinline__ double uniform01(unsigned long long seed, unsigned long long a, unsigned long long b, unsigned c) {
  uint32_t r[4];
  philox4x32_10((uint32_t) a, (uint32_t) (a >> 32), (uint32_t) b, (uint32_t) (b >> 32) ^ c, (uint32_t) seed, (uint32_t) (seed >> 32), r);
  return (u53(r[0], r[1]) + 0.5 / 9007199254740992.0);   // (0, 1)
}
// Gamma(alpha = 0.2, 1) by Marsaglia-Tsang on alpha + 1 and the U^(1/alpha) boost; counter-based, one stream per child
__device__ double gamma02(unsigned long long seed, unsigned long long stream, unsigned long long sub) {
  const double a = 1.2, dd = a - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * dd);
  for (unsigned it = 0; it < 64; ++it) {
    const double u1 = uniform01(seed, stream, sub, 4 * it), u2 = uniform01(seed, stream, sub, 4 * it + 1);
    const double x = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);   // Box-Muller
    const double v0 = 1.0 + cc * x;
    if (v0 <= 0.0) continue;
    const double v = v0 * v0 * v0, u = uniform01(seed, stream, sub, 4 * it + 2);
    if (log(u) < 0.5 * x * x + dd - dd * v + dd * log(v)) {
      const double boost = uniform01(seed, stream, sub, 4 * it + 3);
      return dd * v * pow(boost, 1.0 / 0.2);
    }
  }
  return 1e-12;
}

// tree.cpp:276-283 in the reference's operation order (base = log(..) + 1.25; base *= sqrt(N) / (n + 1); base * prior + value)
__device__ __forceinline__ double puct_score(double pf, double sqrt_n, const Node &c) {
  const double base = __dmul_rn(pf, __ddiv_rn(sqrt_n, (double) (c.visits + 1)));
  const int real = c.visits - c.vloss;
  const double value = real > 0 ? __ddiv_rn(c.value_sum, (double) real) : 0.0;     // Node::value (:74-76)
  return __dadd_rn(__dmul_rn(base, c.prior), value);
}

// Node::selectOptimalMove (tree.cpp:102-116): first child with the strictly greatest score; children over the lanes
__device__ int select_child(const Dev &d, const Node *arena, const Node &n) {
  const int lane = lane_id();
  const int nv = n.visits < d.pf_n ? n.visits : d.pf_n - 1;
  const double pf = d.pf_tbl[nv], sq = sqrt((double) n.visits);
  double best_s = -1.0;
  int best_i = -1;
  for (int i = lane; i < (int) n.n_children; i += 32) {
    const double s = puct_score(pf, sq, arena[n.first_child + i]);
    if (s > best_s) { best_s = s; best_i = i; }
  }
  // warp argmax with "first wins": scores are >= +0.0 (prior, value and the exploration factor are non-negative), so their
  // bit patterns order like unsigned integers: two 32-bit max reductions (REDUX) find the greatest score, a min reduction
  // the smallest child index that has it (five shuffle rounds of a double + an index before)
  if (n.n_children == 0) return -1;
  const unsigned long long key = best_i >= 0 ? (unsigned long long) __double_as_longlong(best_s) : 0ull;
  const unsigned hi = (unsigned) (key >> 32), lo = (unsigned) key;
  const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
  const bool top = best_i >= 0 && hi == mhi && lo == mlo;
  const unsigned first = __reduce_min_sync(0xffffffffu, top ? (unsigned) best_i : 0x7fffffffu);
  return first == 0x7fffffffu ? -1 : n.first_child + (int) first;
}

// Game::tryMove without the mate / stalemate scan (game.cpp:847-871); that scan's outcome is cached in Node::state
__device__ void play_lazy(Pos *pos, int from, int to, int promo) {
  {   // every moved2x flag falls (game.cpp:505-513): two squares per lane instead of 64 trips of one lane
    const int lane = lane_id();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int8_t v = pos->sq[lane + 32 * h];
      if (v == 9) pos->sq[lane + 32 * h] = 8;
      else if (v == -9) pos->sq[lane + 32 * h] = -8;
    }
    __syncwarp();
  }
  if (lane_id() == 0) {
    const bool capture = crules::apply_move_cleared(pos->sq, from, to, promo);
    pos->side = (int8_t) -pos->side;
    pos->no_capture = capture ? 0 : pos->no_capture + 1;
    if (pos->no_capture >= 50) { pos->over = 1; pos->result = 0; }
  }
  __syncwarp();
}

// Game::updateGameState (game.cpp:872-890) for "no legal move": stalemate if the king is safe, else the mover lost
__device__ void mark_over(Pos *pos, Node *n) {
  if (lane_id() == 0) {
    const int ks = crules::king_square(pos->sq, pos->side);
    const bool safe = ks < 0 || !crules::attacked(crules::plain(pos->sq), ks >> 3, ks & 7, pos->side);
    n->state = 2;
    n->result = safe ? 0 : (int8_t) -pos->side;
    pos->over = 1;
    pos->result = n->result;
  }
  __syncwarp();
}

// legal moves of the side to move. Fills ws->moves in the reference's order (from ascending, to ascending; one entry per
// (from, to): the std::map<Move> key, Q8) and returns the count (warp-uniform).
//   (1) pseudo-legal targets with the from-squares over the lanes: lane l owns squares l and l + 32;
//   (2) the (from, to) pairs whose legality needs canPieceMove's simulation (game.cpp:196-226: is the own king attacked
//       with the piece lifted onto `to`?) are flattened into one list and dealt out one PAIR per lane — a queen's twenty
//       simulations no longer run on one lane while thirty-one wait (the per-piece form kept 6.7 lanes busy per
//       instruction in the expand kernel); the verdicts come back as bits OR-ed into per-square masks in shared memory.
// Scratch: the pair list lives in ws->weight (free outside finish_expansion), the masks in ws->child (free until
// make_child); <= 16 pieces x 27 targets = 432 pairs, the list holds 1 024.
__device__ int warp_movegen(WarpShared *ws) {
  const int lane = lane_id();
  const int8_t *sq = ws->pos.sq;
  const int color = ws->pos.side;
  // Board::getKingPosition: first own king in scan order
  const unsigned k0 = __ballot_sync(0xffffffffu, crules::kind_of(sq[lane]) == 1 && crules::color_of(sq[lane]) == color);
  const unsigned k1 = __ballot_sync(0xffffffffu, crules::kind_of(sq[lane + 32]) == 1 && crules::color_of(sq[lane + 32]) == color);
  const int ks = k0 ? __ffs(k0) - 1 : (k1 ? 32 + __ffs(k1) - 1 : -1);
  bool calm = false;
  if (ks >= 0) {
    int c = 0;
    if (lane == 0) c = !crules::attacked(crules::plain(sq), ks >> 3, ks & 7, color);
    calm = __shfl_sync(0xffffffffu, c, 0) != 0;
  }
  uint16_t *pairs = reinterpret_cast<uint16_t *>(ws->weight);
  uint32_t *mask = reinterpret_cast<uint32_t *>(ws->child);   // [64][2]: low / high half of a from-square's legal targets
  uint64_t m[2] = {0ull, 0ull};
  bool sim[2] = {false, false};
#pragma unroll 1                                     // ONE inlined copy of the rules (instruction footprint), two trips
  for (int h = 0; h < 2; ++h) {
    const int from = lane + 32 * h;
    const int k = sq[from];
    const bool own = k != 0 && crules::color_of(k) == color;
    const uint64_t mh = own ? crules::targets(sq, from) : 0ull;
    // a non-king piece that shares no line with a calm king cannot uncover an attack (chess_rules.cuh legal_targets)
    const bool sh = mh != 0 && ks >= 0 && !(calm && crules::kind_of(k) != 1 && !crules::aligned(from, ks));
    mask[2 * from] = 0u; mask[2 * from + 1] = 0u;
    if (h == 0) { m[0] = mh; sim[0] = sh; } else { m[1] = mh; sim[1] = sh; }
  }
  {
    const int p0 = sim[0] ? __popcll(m[0]) : 0, p1 = sim[1] ? __popcll(m[1]) : 0;
    int q = p0 + p1;                                 // this lane's pairs; exclusive prefix over the lanes
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, q, o);
      if (lane >= o) q += t;
    }
    const int n_pairs = __shfl_sync(0xffffffffu, q, 31);
    if (n_pairs > (int) (sizeof(ws->weight) / sizeof(uint16_t))) return -1;   // not a chess position (dozens of queens): the list holds 1 024
    int at = q - p0 - p1;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint64_t mm = sim[h] ? m[h] : 0ull;
      while (mm) {
        const int to = __ffsll((long long) mm) - 1;
        mm &= mm - 1;
        pairs[at++] = (uint16_t) ((lane + 32 * h) | (to << 8));
      }
    }
    __syncwarp();
    for (int i = lane; i < n_pairs; i += 32) {       // Board::canPieceMove for one (from, to) per lane
      const int from = pairs[i] & 0xff, to = pairs[i] >> 8;
      const int piece = sq[from];
      const crules::View v{sq, from, to, piece};
      const int k2 = crules::kind_of(piece) == 1 ? to : ks;   // a king move is tried with the king on `to`
      if (!crules::attacked(v, k2 >> 3, k2 & 7, color)) atomicOr(&mask[2 * from + (to >> 5)], 1u << (to & 31));
    }
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (sim[h]) m[h] = ((uint64_t) mask[2 * (lane + 32 * h) + 1] << 32) | mask[2 * (lane + 32 * h)];
    __syncwarp();                                    // the scratch is free again (make_child writes ws->child)
  }
  const int c0 = __popcll(m[0]), c1 = __popcll(m[1]);
  int s0 = c0, s1 = c1;                              // inclusive prefix sums over the lanes
  for (int o = 1; o < 32; o <<= 1) {
    const int t0 = __shfl_up_sync(0xffffffffu, s0, o), t1 = __shfl_up_sync(0xffffffffu, s1, o);
    if (lane >= o) { s0 += t0; s1 += t1; }
  }
  const int total0 = __shfl_sync(0xffffffffu, s0, 31), total = total0 + __shfl_sync(0xffffffffu, s1, 31);
  if (total > kMaxChildren) return -1;               // cannot happen in chess (218 is the known maximum)
  int at[2] = {s0 - c0, total0 + s1 - c1};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint64_t mm = m[h];
    while (mm) {
      const int to = __ffsll((long long) mm) - 1;
      mm &= mm - 1;
      ws->moves[at[h]++] = (uint16_t) ((lane + 32 * h) | (to << 8));
    }
  }
  __syncwarp();
  return total;
}

__device__ __forceinline__ bool promotes(const int8_t *sq, int from, int to) {
  return crules::kind_of(sq[from]) == 6 && ((to >> 3) == 0 || (to >> 3) == 7);
}

// lane-private child board: the parent with move i made; a promotion becomes the BISHOP's, whose logit is the one the
// reference's std::map keeps for the key (actionMap[move] = ..., decider.cpp:52-56)
// Move::doMove first drops every moved2x flag (game.cpp:505-513) whatever the move: done ONCE per parent, two squares per
// lane, instead of 64 trips per child. Call after warp_movegen (en passant needs the flags) and before make_child.
__device__ __forceinline__ void prepare_children(WarpShared *ws) {
  const int lane = lane_id();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int8_t v = ws->pos.sq[lane + 32 * h];
    ws->base[lane + 32 * h] = v == 9 ? (int8_t) 8 : (v == -9 ? (int8_t) -8 : v);
  }
  __syncwarp();
}
__device__ __forceinline__ int8_t *make_child(WarpShared *ws, int i) {
  int8_t *cb = ws->child + lane_id() * kBoardStride;
  const uint32_t *src = reinterpret_cast<const uint32_t *>(ws->base);
  uint32_t *dst = reinterpret_cast<uint32_t *>(cb);
#pragma unroll
  for (int w = 0; w < 16; ++w) dst[w] = src[w];
  const int from = ws->moves[i] & 0xff, to = ws->moves[i] >> 8;
  crules::apply_move_cleared(cb, from, to, promotes(ws->pos.sq, from, to) ? 7 : 0);
  return cb;
}

// The expansion a parked playout asked for (Decider::prediction's move loop, decider.cpp:43-58, and Node::expand's
// children, tree.cpp:86-95): legal moves, child nodes, the parent and the evaluated child boards into the evaluator's
// input buffer. One warp per request — the requests of a round are independent, so this (the expensive part of a
// playout: move generation and up to 40 child boards) runs at full occupancy while the tree walk stays sequential per tree.
//   no legal move           -> the node is marked finished (mate / stalemate); the playout ends when it resumes
//   evaluation buffer full  -> request dropped; the playout asks again next round
//   arena full              -> the node stays a leaf for good (state 4)
__global__ void __launch_bounds__(kWarpsPerCta * 32) mcts_expand_kernel(const Dev d, int buf) {
  __shared__ WarpShared shared[kWarpsPerCta];
  const int r = (int) (blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5));
  if (r >= *d.req_count) return;
  const int lane = lane_id();
  WarpShared *ws = &shared[threadIdx.x >> 5];
  const int pidx = d.req[r], tree = pidx / d.in_flight;
  Playout *pl = &d.parked[pidx];
  Node *arena = d.nodes + (size_t) tree * d.node_cap;
  const int node = pl->node;
  {
    uint16_t *pd = reinterpret_cast<uint16_t *>(&ws->pos);
    const uint16_t *ps = reinterpret_cast<const uint16_t *>(&pl->pos);
    for (int i = lane; i < (int) (sizeof(Pos) / 2); i += 32) pd[i] = ps[i];
  }
  __syncwarp();
  Node &parent = arena[node];
  const int n = warp_movegen(ws);
  prepare_children(ws);
  if (n <= 0) {                          // Game::updateGameState (game.cpp:872-890): stalemate if the king is safe, else lost
    if (lane == 0) {
      const int ks = crules::king_square(ws->pos.sq, ws->pos.side);
      const bool safe = ks < 0 || !crules::attacked(crules::plain(ws->pos.sq), ks >> 3, ks & 7, ws->pos.side);
      parent.state = 2;
      parent.result = safe ? 0 : (int8_t) -ws->pos.side;
      parent.flags &= (uint8_t) ~kPending;
      pl->slot = -1;
    }
    return;
  }
  // which children go to the network: all of them, or (decider.cpp:52-57 as shipped) those that leave the opponent no reply
  for (int w = lane; w < kMaxChildren / 32; w += 32) ws->eval_mask[w] = 0;
  __syncwarp();
  int n_eval = n;
  if (!d.network_priors) {
    n_eval = 0;
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      bool ask = false;
      if (i < n) ask = !crules::has_legal_move(make_child(ws, i), -ws->pos.side);
      const unsigned b = __ballot_sync(0xffffffffu, ask);
      if (lane == 0) ws->eval_mask[base >> 5] = b;
      n_eval += __popc(b);
    }
  } else {
    for (int w = lane; w < (n + 31) / 32; w += 32) ws->eval_mask[w] = (w + 1) * 32 <= n ? 0xffffffffu : ((1u << (n & 31)) - 1u);
  }
  __syncwarp();
  TreeState *ts = &d.trees[tree];
  int first = 0, s = 0;
  if (lane == 0) {
    first = atomicAdd(&ts->n_nodes, n);
    if ((long) first + n > d.node_cap) {
      atomicSub(&ts->n_nodes, n);
      atomicAdd(&d.cnt->arena_full, 1ull);
      parent.state = 4;
      parent.flags &= (uint8_t) ~kPending;
      pl->slot = -1;
      first = -1;
    } else {
      s = atomicAdd(d.eval_count[buf], 1 + n_eval);
      if (s + 1 + n_eval > d.eval_cap) {
        atomicSub(d.eval_count[buf], 1 + n_eval);
        atomicSub(&ts->n_nodes, n);
        atomicAdd(&d.cnt->eval_full, 1ull);
        parent.flags &= (uint8_t) ~kPending;
        pl->slot = -1;
        first = -1;
      }
    }
  }
  first = __shfl_sync(0xffffffffu, first, 0);
  s = __shfl_sync(0xffffffffu, s, 0);
  if (first < 0) return;
  // the parent board, then the evaluated children in child order
  int8_t *out = d.eval_codes[buf] + (size_t) s * 64;
  reinterpret_cast<uint16_t *>(out)[lane] = reinterpret_cast<const uint16_t *>(ws->pos.sq)[lane];
  int rank = 0;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    const unsigned mask = ws->eval_mask[base >> 5];
    const bool ask = (mask >> lane) & 1u;
    if (i < n && ask) {
      const int8_t *cb = make_child(ws, i);
      uint32_t *dst = reinterpret_cast<uint32_t *>(out + (size_t) (1 + rank + __popc(mask & ((1u << lane) - 1u))) * 64);
      const uint32_t *src = reinterpret_cast<const uint32_t *>(cb);
#pragma unroll
      for (int w = 0; w < 16; ++w) dst[w] = src[w];
    }
    rank += __popc(mask);
  }
  const int8_t child_color = (int8_t) -parent.color;
  for (int i = lane; i < n; i += 32) {
    Node c;
    c.prior = 0.0; c.value_sum = 0.0; c.visits = 0; c.vloss = 0; c.first_child = -1; c.n_children = 0;
    c.color = child_color;
    c.flags = ((ws->eval_mask[i >> 5] >> (i & 31)) & 1u) ? kEval : 0;
    c.state = 0; c.result = 0;
    c.from = (uint8_t) (ws->moves[i] & 0xff); c.to = (uint8_t) (ws->moves[i] >> 8);
    c.promo = promotes(ws->pos.sq, c.from, c.to) ? 3 : 0;
    c.pad[0] = c.pad[1] = c.pad[2] = 0;
    arena[first + i] = c;
  }
  __syncwarp();
  if (lane == 0) {
    parent.first_child = first;
    parent.n_children = (uint16_t) n;
    parent.state = 1;
    pl->slot = s;
    atomicAdd(&d.cnt->expansions, 1ull);
    atomicAdd(&d.cnt->boards, (unsigned long long) (1 + n_eval));
  }
}

// priors of a node's children from the evaluations (MCTS::expand_node tree.cpp:294-306 + Node::expand :86-95), the
// root's Dirichlet noise (:318-347)
__device__ void finish_expansion(const Dev &d, const double *res, int tree, Node *arena, int node, int slot, WarpShared *ws, int move_no) {
  const int lane = lane_id();
  Node &n = arena[node];
  const double cm = n.color;
  const int k = n.n_children, first = n.first_child;
  int rank = 0;
  for (int base = 0; base < k; base += 32) {
    const int i = base + lane;
    bool ev = false;
    if (i < k) ev = arena[first + i].flags & kEval;
    const unsigned mask = __ballot_sync(0xffffffffu, ev);
    if (i < k) {
      // logit = cm * net(child) or cm * 20 (decider.cpp:52-57); weight = exp(cm * logit) (tree.cpp:302), cm * cm = 1
      ws->weight[i] = ev ? exp(cm * (cm * res[slot + 1 + rank + __popc(mask & ((1u << lane) - 1u))])) : d.exp20;
    }
    rank += __popc(mask);
  }
  __syncwarp();
  double sum = 0.0;
  if (lane == 0)
    for (int i = 0; i < k; ++i) sum = __dadd_rn(sum, ws->weight[i]);     // sum_weights in map order (:303)
  sum = __shfl_sync(0xffffffffu, sum, 0);
  for (int i = lane; i < k; i += 32) ws->weight[i] = __ddiv_rn(ws->weight[i], sum);   // Node::expand (:92)
  __syncwarp();
  if (node == 0 && d.noise && k > 0) {
    double tot = 0.0;
    const unsigned long long stream = ((unsigned long long) tree << 20) ^ (unsigned long long) move_no;
    double g[kMaxChildren / 32];
    for (int j = 0, i = lane; i < k; i += 32, ++j) { g[j] = gamma02(d.seed, stream, (unsigned long long) i); tot += g[j]; }
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    for (int j = 0, i = lane; i < k; i += 32, ++j) ws->weight[i] = ws->weight[i] * 0.75 + (g[j] / tot) * 0.25;   // Node::addNoise (:97-101)
    __syncwarp();
  }
  for (int i = lane; i < k; i += 32) arena[first + i].prior = ws->weight[i];
  __syncwarp();
  if (lane == 0) n.flags = (uint8_t) ((n.flags & ~kPending) | kExpanded);
  __syncwarp();
}

struct Walk {                          // the playout in the warp's hands (warp-uniform registers + ws->pos / ws->path)
  int node, depth, path_len, slot;
};

// advances the playout until it parks (false) or completes (true); tree.cpp:202-262
__device__ bool advance(const Dev &d, int tree, Node *arena, WarpShared *ws, Walk &p) {
  const int lane = lane_id();
  Pos *pos = &ws->pos;
  bool truncated = false;
  // The node in hand is read ONCE into registers (one 40-byte round trip to L2 / HBM instead of one per field between the
  // warp barriers); the chosen child's read serves the move to play and is the next iteration's node. Nobody else writes
  // this tree (one warp per tree), and this playout writes only where it parks.
  Node n = arena[p.node];
  while (p.depth < kDepth) {
    if (pos->over) break;
    if (n.state == 2) {
      if (lane == 0) { pos->over = 1; pos->result = n.result; }
      __syncwarp();
      break;
    }
    if (n.state == 4) { truncated = true; break; }        // the arena was full when this node asked: a leaf for good
    if (!(n.flags & kExpanded)) {
      p.slot = -1;
      if (!(n.flags & kPending)) {     // the first playout to get here asks for the expansion (slot -2), later ones just wait
        if (lane == 0) arena[p.node].flags = (uint8_t) (n.flags | kPending);
        p.slot = -2;
      }
      if (lane < p.path_len) { Node &v = arena[ws->path[lane]]; v.visits += 1; v.vloss += 1; }   // virtual loss on the path
      __syncwarp();
      return false;
    }
    const int best = select_child(d, arena, n);
    if (best < 0) break;
    n = arena[best];
    play_lazy(pos, n.from, n.to, n.promo);
    p.node = best;
    if (lane == 0) ws->path[p.path_len] = best;
    ++p.path_len;
    ++p.depth;
    __syncwarp();
  }
  if (!pos->over && !truncated) {      // depth limit: the leaf-value rule needs to know whether the game ended here
    Node &n = arena[p.node];
    if (n.state == 2) {
      if (lane == 0) { pos->over = 1; pos->result = n.result; }
      __syncwarp();
    } else if (n.state == 0) {
      const int nm = warp_movegen(ws);
      if (nm == 0) mark_over(pos, &n);
      else { if (lane == 0) n.state = 1; __syncwarp(); }
    }
  }
  // leaf value (tree.cpp:231-253) and back-propagation (:264-269)
  const double cc = pos->side;
  double value;
  if (pos->over) value = (cc * pos->result + 1.0) / 2.0;
  else {
    int part = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = pos->sq[lane + 32 * h], kd = crules::kind_of(k);
      part += crules::color_of(k) * (kd == 1 ? 100 : (int) ((0x1335900ull >> (4 * kd)) & 15ull));
    }
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    int mat = (int) pos->side * part + 1024;
    mat = mat < 0 ? 0 : (mat > 2048 ? 2048 : mat);
    value = d.sig_tbl[mat];
  }
  if (lane < p.path_len) {
    Node &n = arena[ws->path[lane]];
    n.visits += 1;
    n.value_sum = __dadd_rn(n.value_sum, fabs(cc == (double) n.color ? value : 1.0 - value));
  }
  __syncwarp();
  if (lane == 0) atomicAdd(&d.cnt->playouts, 1ull);
  return true;
}

__device__ void park(const Dev &d, int tree, WarpShared *ws, const Walk &p, int at) {
  Playout *dst = &d.parked[(size_t) tree * d.in_flight + at];
  const int lane = lane_id();
  uint16_t *pd = reinterpret_cast<uint16_t *>(&dst->pos);
  const uint16_t *ps = reinterpret_cast<const uint16_t *>(&ws->pos);
  for (int i = lane; i < (int) (sizeof(Pos) / 2); i += 32) pd[i] = ps[i];
  if (lane < kDepth + 2) dst->path[lane] = ws->path[lane];
  if (lane == 0) {
    dst->node = p.node; dst->depth = p.depth; dst->path_len = p.path_len; dst->slot = p.slot;
    if (p.slot == -2) d.req[atomicAdd(d.req_count, 1)] = tree * d.in_flight + at;    // the expansion kernel's work list
  }
  __syncwarp();
}
__device__ void unpark(const Dev &d, int tree, WarpShared *ws, Walk &p, int at) {
  const Playout *src = &d.parked[(size_t) tree * d.in_flight + at];
  const int lane = lane_id();
  __syncwarp();
  uint16_t *pd = reinterpret_cast<uint16_t *>(&ws->pos);
  const uint16_t *ps = reinterpret_cast<const uint16_t *>(&src->pos);
  for (int i = lane; i < (int) (sizeof(Pos) / 2); i += 32) pd[i] = ps[i];
  if (lane < kDepth + 2) ws->path[lane] = src->path[lane];
  p.node = src->node; p.depth = src->depth; p.path_len = src->path_len; p.slot = src->slot;
  __syncwarp();
}

// One round of one tree: resume what is parked, launch new playouts until `in_flight` are parked or the quota is spent.
__global__ void __launch_bounds__(kWarpsPerCta * 32) mcts_walk_kernel(const Dev d, int buf) {
  __shared__ WarpShared shared[kWarpsPerCta];
  const int tree = (int) (blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5));
  if (tree >= d.n_games * d.roots) return;
  const int lane = lane_id();
  WarpShared *ws = &shared[threadIdx.x >> 5];
  GameState *game = &d.games[tree / d.roots];
  if (game->finished || game->search_done) return;
  TreeState *ts = &d.trees[tree];
  Node *arena = d.nodes + (size_t) tree * d.node_cap;
  const double *res = d.eval_out[buf ^ 1];       // the evaluations the previous round asked for
  int n_parked = ts->n_parked;
  // (1) virtual losses off, expansions finished — for ALL parked playouts before any of them moves again
  for (int i = 0; i < n_parked; ++i) {
    const Playout *pl = &d.parked[(size_t) tree * d.in_flight + i];
    const int path_len = pl->path_len;
    if (lane < path_len) { Node &v = arena[pl->path[lane]]; v.visits -= 1; v.vloss -= 1; }
    __syncwarp();
    if (pl->slot >= 0) finish_expansion(d, res, tree, arena, pl->node, pl->slot, ws, game->moves_played);
  }
  // (2) resume them in order; those that park again are compacted to the front. (3) then new playouts, one at a time
  // until the root exists. One loop, so that advance() is inlined once (instruction footprint).
  int done = ts->done, launched = ts->launched, kept = 0;
  Walk p;
  int budget = d.max_launch;
  for (int i = 0;;) {
    if (i < n_parked) {
      unpark(d, tree, ws, p, i++);
    } else {
      if (!(kept < d.in_flight && launched < d.quota && budget-- > 0)) break;
      if (!(arena[0].flags & kExpanded) && launched > done) break;
      __syncwarp();
      {
        uint16_t *pd = reinterpret_cast<uint16_t *>(&ws->pos);
        const uint16_t *ps = reinterpret_cast<const uint16_t *>(&game->pos);
        for (int w = lane; w < (int) (sizeof(Pos) / 2); w += 32) pd[w] = ps[w];
        if (lane == 0) ws->path[0] = 0;
      }
      __syncwarp();
      p.node = 0; p.depth = 0; p.path_len = 1; p.slot = -1;
      ++launched;
    }
    if (advance(d, tree, arena, ws, p)) ++done;
    else park(d, tree, ws, p, kept++);
  }
  if (lane == 0) { ts->n_parked = kept; ts->done = done; ts->launched = launched; }
}

__device__ void new_root(const Dev &d, int tree, int color) {
  Node r;
  memset(&r, 0, sizeof r);
  r.first_child = -1;
  r.color = (int8_t) color;
  d.nodes[(size_t) tree * d.node_cap] = r;
  d.trees[tree] = TreeState{1, 0, 0, 0};
}

// Per game (one thread each; runs once per move): all roots done -> merge, choose, play, keep a case, new roots.
__global__ void __launch_bounds__(128) mcts_game_kernel(const Dev d) {
  const int g = (int) (blockIdx.x * blockDim.x + threadIdx.x);
  if (g >= d.n_games) return;
  GameState *game = &d.games[g];
  if (game->finished || game->search_done) return;
  for (int r = 0; r < d.roots; ++r) {
    const TreeState &ts = d.trees[g * d.roots + r];
    if (ts.done < d.quota || ts.n_parked > 0) return;
  }
  if (d.stop_after_search) {
    game->search_done = 1;
    atomicAdd(&d.cnt->searches_done, 1ull);
    return;
  }
  // Node::combineNodes (tree.cpp:122-142): the roots saw the same position, so their children are the same moves in the
  // same order; priors are averaged, visit counts and value sums added, for the children and for the root
  const Node *root0 = d.nodes + (size_t) (g * d.roots) * d.node_cap;
  int root_visits = 0;
  double root_value_sum = 0.0;
  for (int r = 0; r < d.roots; ++r) {
    const Node &rt = d.nodes[(size_t) (g * d.roots + r) * d.node_cap];
    root_visits += rt.visits;
    root_value_sum += rt.value_sum;
  }
  const int k = root0->n_children;
  const int nv = root_visits < d.pf_n ? root_visits : d.pf_n - 1;
  const double pf = d.pf_tbl[nv], sq = sqrt((double) root_visits);
  int best = -1;
  double max_score = -1.0;
  for (int i = 0; i < k; ++i) {
    Node c;
    c.prior = 0.0; c.value_sum = 0.0; c.visits = 0; c.vloss = 0;
    for (int r = 0; r < d.roots; ++r) {
      const Node *arena = d.nodes + (size_t) (g * d.roots + r) * d.node_cap;
      const Node &rc = arena[arena[0].first_child + i];
      c.prior += rc.prior / d.roots;
      c.visits += rc.visits;
      c.value_sum += rc.value_sum;
    }
    const double s = puct_score(pf, sq, c);
    if (s > max_score) { max_score = s; best = i; }
  }
  Pos *pos = &game->pos;
  if (best >= 0) {                       // Game::tryMove (game.cpp:847-871) for a move known to be legal
    const Node &mv = root0[root0->first_child + best];
    const bool capture = crules::apply_move(pos->sq, mv.from, mv.to, mv.promo);
    pos->side = (int8_t) -pos->side;
    pos->no_capture = capture ? 0 : pos->no_capture + 1;
    if (pos->no_capture >= 50) { pos->over = 1; pos->result = 0; }
    else if (!crules::has_legal_move(pos->sq, pos->side)) {   // Game::updateGameState (:872-890)
      const int ks = crules::king_square(pos->sq, pos->side);
      const bool safe = ks < 0 || !crules::attacked(crules::plain(pos->sq), ks >> 3, ks & 7, pos->side);
      pos->over = 1;
      pos->result = safe ? 0 : (int8_t) -pos->side;
    }
  }
  game->moves_played += 1;
  atomicAdd(&d.cnt->moves, 1ull);
  // NetworkStorage::saveBoard -> SIMULATION_BOARD_SAVE_PROCEDURE (network.cpp:372-378, initialization.cpp:221-227)
  if (d.save_ratio > 0.0 && uniform01(d.seed ^ 0x5ca1ab1eull, (unsigned long long) g, (unsigned long long) game->moves_played, 7u) < d.save_ratio) {
    const unsigned long long at = atomicAdd(&d.cnt->cases, 1ull);
    if (at < (unsigned long long) d.cases_cap) {
      KeptCase &kc = d.cases[at];
      for (int i = 0; i < 64; ++i) kc.codes[i] = pos->sq[i];
      const double v = root_visits > 0 ? root_value_sum / root_visits : 0.0;
      const double b = 2.0 * v - 1.0;
      kc.bounded = b < -1.0 ? -1.0 : (b > 1.0 ? 1.0 : b);
      kc.game = g;
      kc.pad = 0;
    }
  }
  if (best < 0 || pos->over || game->moves_played >= d.max_moves) {
    game->finished = 1;
    atomicAdd(&d.cnt->games_finished, 1ull);
  } else {
    for (int r = 0; r < d.roots; ++r) new_root(d, g * d.roots + r, pos->side);
  }
}

// a new cab_mcts_run: the "searched, waiting to be read" marks of a previous search-only call fall (the finished trees are
// either re-marked at once, search-only again, or their move is played)
__global__ void __launch_bounds__(128) mcts_clear_search_kernel(const Dev d) {
  const int g = (int) (blockIdx.x * blockDim.x + threadIdx.x);
  if (g == 0) d.cnt->searches_done = 0ull;
  if (g < d.n_games) d.games[g].search_done = 0;
}

__global__ void __launch_bounds__(128) mcts_reset_kernel(const Dev d) {
  const int g = (int) (blockIdx.x * blockDim.x + threadIdx.x);
  if (g >= d.n_games) return;
  d.games[g].moves_played = 0;
  d.games[g].finished = 0;
  d.games[g].search_done = 0;
  for (int r = 0; r < d.roots; ++r) new_root(d, g * d.roots + r, d.games[g].pos.side);
}

// legal move lists of n positions (tests: the device rules against the recorded reference games); one warp per position
__global__ void __launch_bounds__(kWarpsPerCta * 32) movegen_kernel(const int8_t *codes, const int8_t *side, int n, int queen_only,
                                                                    int *n_moves, uint8_t *moves, int cap, uint8_t *flags) {
  __shared__ WarpShared shared[kWarpsPerCta];
  const int p = (int) (blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5));
  if (p >= n) return;
  const int lane = lane_id();
  WarpShared *ws = &shared[threadIdx.x >> 5];
  reinterpret_cast<uint16_t *>(ws->pos.sq)[lane] = reinterpret_cast<const uint16_t *>(codes + (size_t) p * 64)[lane];
  if (lane == 0) { ws->pos.side = side[p]; ws->pos.over = 0; ws->pos.result = 0; ws->pos.no_capture = 0; }
  __syncwarp();
  const int k = warp_movegen(ws);
  prepare_children(ws);
  // expand promotions to Q R N B unless queen_only (game.cpp:404-417)
  int total = 0;
  if (lane == 0) {
    uint8_t *out = moves + (size_t) p * cap * 3;
    for (int i = 0; i < k; ++i) {
      const int from = ws->moves[i] & 0xff, to = ws->moves[i] >> 8;
      const bool pr = promotes(ws->pos.sq, from, to);
      const int reps = pr && !queen_only ? 4 : 1;
      for (int r = 0; r < reps; ++r, ++total)
        if (total < cap) { out[3 * total] = (uint8_t) from; out[3 * total + 1] = (uint8_t) to; out[3 * total + 2] = (uint8_t) (pr ? (int) ((0x7643u >> (4 * r)) & 15u) : 0); }
    }
    n_moves[p] = total;
  }
  // per queen-only child: does the move leave the opponent without a reply? (decider.cpp:47-52)
  if (flags != nullptr)
    for (int base = 0; base < k; base += 32) {
      const int i = base + lane;
      if (i < k && i < cap) flags[(size_t) p * cap + i] = crules::has_legal_move(make_child(ws, i), -ws->pos.side) ? 0 : 1;
    }
}

}  // namespace mcts
}  // namespace cab

using namespace cab;
using namespace cab::mcts;

struct cab_mcts {
  cab_net *net = nullptr;
  cab_mcts_params prm{};
  Dev d{};
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t eval_stream[4] = {};      // a round's boards are evaluated in pieces over these (fork / join on `stream`)
  cudaEvent_t ev_fork = nullptr, ev_join[4] = {};
  double *d_pf = nullptr, *d_sig = nullptr;
  int *h_count = nullptr;         // pinned: boards queued by the last round
  long rounds = 0;
  double eval_ms = 0.0, wall_s = 0.0;
  int round_parity = 0;
  bool positions_set = false;
};

static void mcts_free(cab_mcts *q) {
  if (!q) return;
  cudaFree(q->d.req); cudaFree(q->d.req_count);
  cudaFree(q->d.nodes); cudaFree(q->d.trees); cudaFree(q->d.parked); cudaFree(q->d.games); cudaFree(q->d.cnt); cudaFree(q->d.cases);
  for (int b = 0; b < 2; ++b) { cudaFree(q->d.eval_codes[b]); cudaFree(q->d.eval_out[b]); cudaFree(q->d.eval_count[b]); }
  cudaFree(q->d_pf); cudaFree(q->d_sig);
  if (q->h_count) cudaFreeHost(q->h_count);
  if (q->stream) cudaStreamDestroy(q->stream);
  for (auto st : q->eval_stream) if (st) cudaStreamDestroy(st);
  for (auto e : q->ev_join) if (e) cudaEventDestroy(e);
  if (q->ev_fork) cudaEventDestroy(q->ev_fork);
  if (q->ev0) cudaEventDestroy(q->ev0);
  if (q->ev1) cudaEventDestroy(q->ev1);
  delete q;
}

static int mcts_build(cab_mcts *q) {
  const cab_mcts_params &p = q->prm;
  Dev &d = q->d;
  d.n_games = p.games; d.roots = p.roots; d.quota = std::max(1, p.playouts_per_move / p.roots);
  d.in_flight = p.in_flight; d.max_moves = p.max_moves; d.network_priors = p.network_priors; d.noise = p.noise;
  d.stop_after_search = 0; d.max_launch = std::max(4 * p.in_flight, 8);
  d.save_ratio = p.save_ratio; d.seed = p.seed;
  const long trees = (long) p.games * p.roots;
  d.node_cap = 1 + (long) d.quota * 320;                  // up to 8 expansions of ~40 children per playout
  CAB_CUDA(cudaMalloc(&d.nodes, (size_t) trees * d.node_cap * sizeof(Node)));
  CAB_CUDA(cudaMalloc(&d.trees, trees * sizeof(TreeState)));
  CAB_CUDA(cudaMalloc(&d.parked, (size_t) trees * p.in_flight * sizeof(Playout)));
  CAB_CUDA(cudaMalloc(&d.games, p.games * sizeof(GameState)));
  CAB_CUDA(cudaMalloc(&d.req, (size_t) trees * p.in_flight * sizeof(int)));
  CAB_CUDA(cudaMalloc(&d.req_count, sizeof(int)));
  CAB_CUDA(memset_now(d.req_count, 0, sizeof(int)));
  CAB_CUDA(cudaMalloc(&d.cnt, sizeof(Counters)));
  CAB_CUDA(memset_now(d.cnt, 0, sizeof(Counters)));
  d.cases_cap = p.save_ratio > 0.0 ? (int) std::min<long>((long) p.games * (p.max_moves + 1), 1L << 24) : 1;
  CAB_CUDA(cudaMalloc(&d.cases, (size_t) d.cases_cap * sizeof(KeptCase)));
  // one round queues at most trees * in_flight expansions; ~40 boards each when every child is evaluated
  const long want = trees * p.in_flight * (p.network_priors ? 48L : 4L) + 4096;
  d.eval_cap = (int) std::min<long>(want, 1L << 24);
  for (int b = 0; b < 2; ++b) {
    CAB_CUDA(cudaMalloc(&d.eval_codes[b], (size_t) d.eval_cap * 64));
    CAB_CUDA(cudaMalloc(&d.eval_out[b], (size_t) d.eval_cap * sizeof(double)));
    CAB_CUDA(cudaMalloc(&d.eval_count[b], sizeof(int)));
    CAB_CUDA(memset_now(d.eval_count[b], 0, sizeof(int)));
  }
  // tables from the host's libm: the very values the reference's search computes
  d.pf_n = d.quota * p.roots + p.in_flight * p.roots + 64;
  std::vector<double> pf(d.pf_n), sig(2049);
  for (int n = 0; n < d.pf_n; ++n) pf[n] = std::log((n + 19652.0 + 1.0) / 19652.0) + 1.25;      // tree.cpp:280
  for (int m = 0; m <= 2048; ++m) { const double value = (double) (m - 1024); sig[m] = 1.0 / (1.0 + std::exp(-value / 27.18)); }  // :252
  CAB_CUDA(cudaMalloc(&q->d_pf, pf.size() * sizeof(double)));
  CAB_CUDA(cudaMalloc(&q->d_sig, sig.size() * sizeof(double)));
  CAB_CUDA(upload_now(q->d_pf, pf.data(), pf.size() * sizeof(double)));
  CAB_CUDA(upload_now(q->d_sig, sig.data(), sig.size() * sizeof(double)));
  d.pf_tbl = q->d_pf; d.sig_tbl = q->d_sig;
  d.exp20 = std::exp(20.0);
  CAB_CUDA(cudaMallocHost(&q->h_count, sizeof(int)));
  CAB_CUDA(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking));
  CAB_CUDA(cudaEventCreate(&q->ev0));
  CAB_CUDA(cudaEventCreate(&q->ev1));
  CAB_CUDA(cudaEventCreateWithFlags(&q->ev_fork, cudaEventDisableTiming));
  for (int i = 0; i < 4; ++i) {
    CAB_CUDA(cudaStreamCreateWithFlags(&q->eval_stream[i], cudaStreamNonBlocking));
    CAB_CUDA(cudaEventCreateWithFlags(&q->ev_join[i], cudaEventDisableTiming));
  }
  // default: every game from the start position
  std::vector<GameState> gs(p.games);
  static const int8_t back[8] = {4, 6, 7, 3, 1, 7, 6, 4};
  for (auto &g : gs) {
    memset(&g, 0, sizeof g);
    for (int c = 0; c < 8; ++c) { g.pos.sq[c] = back[c]; g.pos.sq[8 + c] = 8; g.pos.sq[48 + c] = -8; g.pos.sq[56 + c] = (int8_t) -back[c]; }
    g.pos.side = 1;
  }
  CAB_CUDA(upload_now(d.games, gs.data(), gs.size() * sizeof(GameState)));
  mcts_reset_kernel<<<(p.games + 127) / 128, 128, 0, q->stream>>>(d);
  CAB_CUDA(cudaGetLastError());
  CAB_CUDA(cudaStreamSynchronize(q->stream));
  return CAB_OK;
}

extern "C" {

int cab_mcts_create(cab_net *net, const cab_mcts_params *params, cab_mcts **out) {
  if (!net || !params || !out) { set_error("null argument"); return CAB_EINVAL; }
  *out = nullptr;
  const cab_mcts_params &p = *params;
  if (p.games <= 0 || p.roots <= 0 || p.roots > 64 || p.playouts_per_move <= 0 || p.in_flight <= 0 || p.in_flight > 4096 || p.max_moves < 0 ||
      p.save_ratio < 0.0 || p.save_ratio > 1.0 || p.precision < 0 || p.precision > 1) { set_error("bad MCTS parameters"); return CAB_EINVAL; }
  if (!net->fp64_ok) { set_error("the network does not fit the fp64 evaluator"); return CAB_EINVAL; }
  CAB_CUDA(cudaSetDevice(net->device));
  cab_mcts *q = new cab_mcts();
  q->net = net;
  q->prm = p;
  const int rc = mcts_build(q);
  if (rc != CAB_OK) { mcts_free(q); return rc; }
  *out = q;
  return CAB_OK;
}

int cab_mcts_destroy(cab_mcts *q) {
  if (!q) return CAB_OK;
  cudaSetDevice(q->net->device);
  mcts_free(q);
  return CAB_OK;
}

// positions the games start from (default: the start position): codes [games][64], side [games] (+1 / -1),
// no_capture [games] (plies since the last capture; may be null = 0). Restarts every game.
int cab_mcts_set_positions(cab_mcts *q, const int8_t *codes, const int8_t *side, const int *no_capture) {
  if (!q || !codes || !side) { set_error("null argument"); return CAB_EINVAL; }
  CAB_CUDA(cudaSetDevice(q->net->device));
  std::vector<GameState> gs(q->prm.games);
  for (int g = 0; g < q->prm.games; ++g) {
    memset(&gs[g], 0, sizeof(GameState));
    memcpy(gs[g].pos.sq, codes + (size_t) g * 64, 64);
    if (side[g] != 1 && side[g] != -1) { set_error("side must be +1 or -1"); return CAB_EINVAL; }
    gs[g].pos.side = side[g];
    gs[g].pos.no_capture = no_capture ? no_capture[g] : 0;
  }
  CAB_CUDA(upload_now(q->d.games, gs.data(), gs.size() * sizeof(GameState)));
  CAB_CUDA(memset_now(q->d.cnt, 0, sizeof(Counters)));
  for (int b = 0; b < 2; ++b) CAB_CUDA(memset_now(q->d.eval_count[b], 0, sizeof(int)));
  mcts_reset_kernel<<<(q->prm.games + 127) / 128, 128, 0, q->stream>>>(q->d);
  CAB_CUDA(cudaGetLastError());
  CAB_CUDA(cudaStreamSynchronize(q->stream));
  q->rounds = 0; q->eval_ms = 0.0; q->wall_s = 0.0; q->round_parity = 0;
  return CAB_OK;
}

// Runs rounds until every game is finished (search_only = 0) or until every game's CURRENT move has been searched
// (search_only = 1: nothing is played; read the roots with cab_mcts_root_children). max_rounds > 0 bounds the call.
int cab_mcts_run(cab_mcts *q, int search_only, long max_rounds, cab_mcts_stats *stats) {
  if (!q) { set_error("null argument"); return CAB_EINVAL; }
  cab_net *net = q->net;
  CAB_CUDA(cudaSetDevice(net->device));
  {
    std::lock_guard<std::mutex> lk(net->mu);
    const int rc = ensure_packed(net, q->stream, true);
    if (rc != CAB_OK) return rc;
  }
  Dev d = q->d;
  d.stop_after_search = search_only ? 1 : 0;
  const int n_trees = d.n_games * d.roots;
  const auto t0 = std::chrono::steady_clock::now();
  Counters h{};
  long rounds = 0, idle_rounds = 0;
  unsigned long long last_playouts = ~0ull, last_moves = ~0ull;
  mcts_clear_search_kernel<<<(d.n_games + 127) / 128, 128, 0, q->stream>>>(d);
  for (;;) {
    const int buf = q->round_parity;
    mcts_game_kernel<<<(d.n_games + 127) / 128, 128, 0, q->stream>>>(d);
    CAB_CUDA(cudaMemsetAsync(d.eval_count[buf], 0, sizeof(int), q->stream));
    CAB_CUDA(cudaMemsetAsync(d.req_count, 0, sizeof(int), q->stream));
    mcts_walk_kernel<<<(n_trees + kWarpsPerCta - 1) / kWarpsPerCta, kWarpsPerCta * 32, 0, q->stream>>>(d, buf);
    // one warp per request; the grid covers the most a round can ask for, warps beyond the count leave at once
    mcts_expand_kernel<<<(n_trees * d.in_flight + kWarpsPerCta - 1) / kWarpsPerCta, kWarpsPerCta * 32, 0, q->stream>>>(d, buf);
    CAB_CUDA(cudaGetLastError());
    CAB_CUDA(cudaMemcpyAsync(q->h_count, d.eval_count[buf], sizeof(int), cudaMemcpyDeviceToHost, q->stream));
    CAB_CUDA(cudaMemcpyAsync(&h, d.cnt, sizeof(Counters), cudaMemcpyDeviceToHost, q->stream));
    CAB_CUDA(cudaStreamSynchronize(q->stream));
    const int n_boards = *q->h_count;
    net->launches += 3;
    ++rounds;
    q->round_parity ^= 1;
    if (n_boards > 0) {
      CAB_CUDA(cudaEventRecord(q->ev0, q->stream));
      // precision 1: the fused fp16 tcgen05 evaluator (reduced precision: the search is then NOT the reference's, see the header)
      int rc;
      if (q->prm.precision == 1) {
        rc = cab_eval_f16_dev(net, d.eval_codes[buf], n_boards, d.eval_out[buf], q->stream);
      } else if (n_boards <= 2 * kEvalPiece) {
        rc = launch_forward(net, d.eval_codes[buf], n_boards, d.eval_out[buf], nullptr, 0, q->stream, nullptr, /*alone=*/true);
      } else {
        // a big round as pieces over four streams: one persistent launch has every CTA ask for the same weight chunk at the
        // same moment (1.62e8 boards/s), desynchronised pieces do not (1.95e8)
        CAB_CUDA(cudaEventRecord(q->ev_fork, q->stream));
        for (auto st : q->eval_stream) CAB_CUDA(cudaStreamWaitEvent(st, q->ev_fork, 0));
        rc = cab_eval_dev_batches(net, d.eval_codes[buf], n_boards, kEvalPiece, d.eval_out[buf], reinterpret_cast<void *const *>(q->eval_stream), 4);
        for (int i = 0; i < 4 && rc == CAB_OK; ++i) {
          CAB_CUDA(cudaEventRecord(q->ev_join[i], q->eval_stream[i]));
          CAB_CUDA(cudaStreamWaitEvent(q->stream, q->ev_join[i], 0));
        }
      }
      if (rc != CAB_OK) return rc;
      CAB_CUDA(cudaEventRecord(q->ev1, q->stream));
      CAB_CUDA(cudaEventSynchronize(q->ev1));
      float ms = 0.f;
      CAB_CUDA(cudaEventElapsedTime(&ms, q->ev0, q->ev1));
      q->eval_ms += ms;
    }
    const unsigned long long done = search_only ? h.searches_done : h.games_finished;
    if (done >= (unsigned long long) d.n_games && n_boards == 0) break;
    if (max_rounds > 0 && rounds >= max_rounds) break;
    // a round that neither evaluates, nor completes a playout, nor plays a move cannot be followed by one that does
    idle_rounds = (n_boards == 0 && h.playouts == last_playouts && h.moves == last_moves) ? idle_rounds + 1 : 0;
    last_playouts = h.playouts; last_moves = h.moves;
    if (idle_rounds >= 16) { set_error("MCTS: %ld rounds without progress (%llu of %d games done)", idle_rounds, done, d.n_games); return CAB_ESTATE; }
  }
  q->rounds += rounds;
  q->wall_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (stats) {
    CAB_CUDA(cudaMemcpy(&h, d.cnt, sizeof(Counters), cudaMemcpyDeviceToHost));
    stats->playouts = (long) h.playouts; stats->expansions = (long) h.expansions; stats->boards_evaluated = (long) h.boards;
    stats->moves_played = (long) h.moves; stats->games_finished = (long) h.games_finished; stats->searches_done = (long) h.searches_done;
    stats->rounds = q->rounds; stats->arena_full = (long) h.arena_full; stats->eval_full = (long) h.eval_full;
    stats->cases_kept = (long) std::min<unsigned long long>(h.cases, (unsigned long long) d.cases_cap);
    stats->seconds = q->wall_s; stats->evaluator_seconds = q->eval_ms * 1e-3;
  }
  return CAB_OK;
}

// The searched root of one game, roots merged like Node::combineNodes (tree.cpp:122-142): per child (map order) from,
// to, prior (mean over roots), visits, value_sum; *root_visits / *root_value_sum for the root itself. root = -1 merges,
// root >= 0 returns that tree alone. Returns the number of children in *n_children (may exceed cap: truncated).
int cab_mcts_root_children(cab_mcts *q, int game, int root, int cap, int *from, int *to, double *prior, int *visits,
                           double *value_sum, int *n_children, int *root_visits, double *root_value_sum) {
  if (!q || game < 0 || game >= q->prm.games || root < -1 || root >= q->prm.roots || !n_children) { set_error("bad argument"); return CAB_EINVAL; }
  CAB_CUDA(cudaSetDevice(q->net->device));
  const Dev &d = q->d;
  const int r0 = root < 0 ? 0 : root, r1 = root < 0 ? q->prm.roots : root + 1;
  int k = 0, rv = 0;
  double rvs = 0.0;
  std::vector<Node> kids;
  std::vector<double> pr; std::vector<int> vi; std::vector<double> vs; std::vector<int> fr, tt;
  for (int r = r0; r < r1; ++r) {
    const Node *arena = d.nodes + (size_t) (game * q->prm.roots + r) * d.node_cap;
    Node rt;
    CAB_CUDA(cudaMemcpy(&rt, arena, sizeof(Node), cudaMemcpyDeviceToHost));
    rv += rt.visits; rvs += rt.value_sum;
    if (rt.first_child < 0) continue;
    kids.resize(rt.n_children);
    CAB_CUDA(cudaMemcpy(kids.data(), arena + rt.first_child, kids.size() * sizeof(Node), cudaMemcpyDeviceToHost));
    if (r == r0) {
      k = rt.n_children;
      pr.assign(k, 0.0); vi.assign(k, 0); vs.assign(k, 0.0); fr.resize(k); tt.resize(k);
      for (int i = 0; i < k; ++i) { fr[i] = kids[i].from; tt[i] = kids[i].to; }
    }
    for (int i = 0; i < k && i < (int) kids.size(); ++i) { pr[i] += kids[i].prior / (r1 - r0); vi[i] += kids[i].visits; vs[i] += kids[i].value_sum; }
  }
  *n_children = k;
  for (int i = 0; i < k && i < cap; ++i) {
    if (from) from[i] = fr[i];
    if (to) to[i] = tt[i];
    if (prior) prior[i] = pr[i];
    if (visits) visits[i] = vi[i];
    if (value_sum) value_sum[i] = vs[i];
  }
  if (root_visits) *root_visits = rv;
  if (root_value_sum) *root_value_sum = rvs;
  return CAB_OK;
}

// current position of a game: codes64 (64 bytes), side, over, result, moves played
int cab_mcts_game_state(cab_mcts *q, int game, int8_t *codes64, int *side, int *over, int *result, int *moves_played) {
  if (!q || game < 0 || game >= q->prm.games) { set_error("bad argument"); return CAB_EINVAL; }
  CAB_CUDA(cudaSetDevice(q->net->device));
  GameState g;
  CAB_CUDA(cudaMemcpy(&g, q->d.games + game, sizeof g, cudaMemcpyDeviceToHost));
  if (codes64) memcpy(codes64, g.pos.sq, 64);
  if (side) *side = g.pos.side;
  if (over) *over = g.pos.over;
  if (result) *result = g.pos.result;
  if (moves_played) *moves_played = g.moves_played;
  return CAB_OK;
}

// Training cases of the FINISHED games: boards kept with probability save_ratio and their targets
// 10 * (4 * clamp(2v - 1) + result) / 5 (make_cases.cpp:36-42). codes_out [cap][64], targets_out [cap]; *n = count.
int cab_mcts_cases(cab_mcts *q, int8_t *codes_out, double *targets_out, long cap, long *n) {
  if (!q || !n) { set_error("null argument"); return CAB_EINVAL; }
  CAB_CUDA(cudaSetDevice(q->net->device));
  Counters h;
  CAB_CUDA(cudaMemcpy(&h, q->d.cnt, sizeof h, cudaMemcpyDeviceToHost));
  const long kept = (long) std::min<unsigned long long>(h.cases, (unsigned long long) q->d.cases_cap);
  std::vector<KeptCase> kc(kept);
  std::vector<GameState> gs(q->prm.games);
  if (kept) CAB_CUDA(cudaMemcpy(kc.data(), q->d.cases, kept * sizeof(KeptCase), cudaMemcpyDeviceToHost));
  CAB_CUDA(cudaMemcpy(gs.data(), q->d.games, gs.size() * sizeof(GameState), cudaMemcpyDeviceToHost));
  long out = 0;
  for (long i = 0; i < kept; ++i) {
    const GameState &g = gs[kc[i].game];
    if (!g.finished) continue;
    if (out < cap && codes_out && targets_out) {
      memcpy(codes_out + out * 64, kc[i].codes, 64);
      const double result = g.pos.over ? (double) g.pos.result : 0.0;          // GameResult::evaluate, game.fwd.h:82-95
      targets_out[out] = 10.0 * (4.0 * kc[i].bounded + 1.0 * result) / 5.0;    // get_overall_result, make_cases.cpp:36-42
    }
    ++out;
  }
  *n = out;
  return CAB_OK;
}

// Legal moves of n positions by the DEVICE rules (the kernels' own move generator), in the reference's order:
// moves_out [n][cap][3] = (from, to, promo), n_moves_out [n]; no_reply_out (optional) [n][cap]: for the i-th
// queen-only move, 1 when it leaves the opponent without a legal move (decider.cpp:47-52; needs queen_only = 1).
int cab_movegen_host(int device, const int8_t *codes, const int8_t *side, int n, int queen_only, int cap, int *n_moves_out,
                     uint8_t *moves_out, uint8_t *no_reply_out) {
  if (n < 0 || cap <= 0 || (n > 0 && (!codes || !side || !n_moves_out || !moves_out))) { set_error("bad argument"); return CAB_EINVAL; }
  if (no_reply_out && !queen_only) { set_error("no_reply_out needs queen_only = 1"); return CAB_EINVAL; }
  if (n == 0) return CAB_OK;
  CAB_CUDA(cudaSetDevice(device));
  struct Dev1 { void *p = nullptr; ~Dev1() { cudaFree(p); } } dc, ds, dn, dm, df;
  CAB_CUDA(cudaMalloc(&dc.p, (size_t) n * 64));
  CAB_CUDA(cudaMalloc(&ds.p, n));
  CAB_CUDA(cudaMalloc(&dn.p, n * sizeof(int)));
  CAB_CUDA(cudaMalloc(&dm.p, (size_t) n * cap * 3));
  if (no_reply_out) { CAB_CUDA(cudaMalloc(&df.p, (size_t) n * cap)); CAB_CUDA(memset_now(df.p, 0, (size_t) n * cap)); }
  CAB_CUDA(upload_now(dc.p, codes, (size_t) n * 64));
  CAB_CUDA(upload_now(ds.p, side, n));
  movegen_kernel<<<(n + kWarpsPerCta - 1) / kWarpsPerCta, kWarpsPerCta * 32>>>((const int8_t *) dc.p, (const int8_t *) ds.p, n, queen_only,
                                                                                (int *) dn.p, (uint8_t *) dm.p, cap, (uint8_t *) df.p);
  CAB_CUDA(cudaGetLastError());
  CAB_CUDA(cudaMemcpy(n_moves_out, dn.p, n * sizeof(int), cudaMemcpyDeviceToHost));
  CAB_CUDA(cudaMemcpy(moves_out, dm.p, (size_t) n * cap * 3, cudaMemcpyDeviceToHost));
  if (no_reply_out) CAB_CUDA(cudaMemcpy(no_reply_out, df.p, (size_t) n * cap, cudaMemcpyDeviceToHost));
  return CAB_OK;
}

}  // extern "C"
