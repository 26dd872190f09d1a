// ref_harness.cpp — C-ABI wrapper around the UNMODIFIED reference classes, compiled together with the
// reference's own translation units where they lie under /root/reference (see oracle/Makefile).
//
// TEST INFRASTRUCTURE ONLY. Output goes to oracle/_ref/libchessai_ref.so. Only tests/, bench.py's
// cpu_baseline / --impl reference leg and __graft_entry__.smoke() may load it; the product path never does.
//
// Everything here calls the reference's own code:
//   network::Network::predictPosition        /root/reference/src/mcts_network/network.cpp:92-117
//   network::Optimizer::optimize / dropout    network.cpp:223-292
//   operator>> / operator<< (Network*&)       network.cpp:386-426
//   game::operator>>(istream&, Board*&)       /root/reference/src/chess/game.cpp:343-370
//   piece::Piece::code() and overrides        /root/reference/src/chess/piece.cpp:103-105,145-147,179-181,260-262
//   network::load_training_case               /root/reference/src/main/network/make_cases.cpp:123-129
//   decider::Decider::prediction              /root/reference/src/mcts_network/decider.cpp:31-61
// Compiled with -fno-access-control so private members (predictPosition, _connections) are reachable
// without touching the reference sources.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cmath>
#include <cstring>
#include <random>
#include <fstream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "mcts_network/network.h"
#include "mcts_network/decider.h"
#include "mcts_network/tree.h"
#include "chess/game.h"
#include "chess/piece.h"
#include "main/network/make_cases.h"
#include "util/math_util.h"
#include "util/string_util.h"

namespace {

// code k in [-9, 9] -> the piece line of the reference's board text format (game.cpp:372-379, piece.cpp:300-375)
const char *piece_text(int k) {
  switch (k) {
    case 0: return "none none .";
    case 1: return "king white false";
    case 2: return "king white true";
    case 3: return "queen white .";
    case 4: return "rook white false";
    case 5: return "rook white true";
    case 6: return "knight white .";
    case 7: return "bishop white .";
    case 8: return "pawn white false";
    case 9: return "pawn white true";
    case -1: return "king black false";
    case -2: return "king black true";
    case -3: return "queen black .";
    case -4: return "rook black false";
    case -5: return "rook black true";
    case -6: return "knight black .";
    case -7: return "bishop black .";
    case -8: return "pawn black false";
    case -9: return "pawn black true";
    default: return nullptr;
  }
}

game::Board *board_from_codes(const int8_t *codes) {
  std::stringstream ss;
  ss << "8 8\n";
  for (int i = 0; i < 64; ++i) {
    const char *txt = piece_text(codes[i]);
    if (txt == nullptr) return nullptr;
    ss << i << " - " << txt << "\n";
  }
  auto *b = new game::Board(8, 8);
  ss >> b;  // the reference's own parser
  return b;
}

}  // namespace

extern "C" {

// ---- network lifetime / IO -------------------------------------------------------------------------
void *ref_net_load(const char *path) { return network::Network::loadFromFile(path); }

void *ref_net_new(const int *dims, int n_dims) {
  std::vector<int> d(dims, dims + n_dims);
  return new network::Network(d);
}

// Network::randomNetwork(dims) — draws from the reference's own (default-seeded, per-TU) mt19937.
void *ref_net_random(const int *dims, int n_dims) {
  std::vector<int> d(dims, dims + n_dims);
  return network::Network::randomNetwork(d);
}

void *ref_net_clone(void *net) { return static_cast<network::Network *>(net)->clone(); }

void ref_net_free(void *net) { delete static_cast<network::Network *>(net); }

int ref_net_num_layers(void *net) { return static_cast<network::Network *>(net)->_num_layers; }

void ref_net_dims(void *net, int *out) {
  auto *n = static_cast<network::Network *>(net);
  for (int i = 0; i < n->_num_layers; ++i) out[i] = n->_dimensions[i];
}

long ref_net_num_weights(void *net) {
  auto *n = static_cast<network::Network *>(net);
  long c = 0;
  for (int l = 0; l + 1 < n->_num_layers; ++l) c += (long) n->_dimensions[l] * n->_dimensions[l + 1];
  return c;
}

// flat n,i,j order == line 3 of the dump format
void ref_net_get_weights(void *net, double *out) {
  auto *n = static_cast<network::Network *>(net);
  long c = 0;
  for (int l = 0; l + 1 < n->_num_layers; ++l)
    for (int i = 0; i < n->_dimensions[l]; ++i)
      for (int j = 0; j < n->_dimensions[l + 1]; ++j) out[c++] = n->_connections[l][i][j];
}

void ref_net_set_weights(void *net, const double *in) {
  auto *n = static_cast<network::Network *>(net);
  long c = 0;
  for (int l = 0; l + 1 < n->_num_layers; ++l)
    for (int i = 0; i < n->_dimensions[l]; ++i)
      for (int j = 0; j < n->_dimensions[l + 1]; ++j) n->_connections[l][i][j] = in[c++];
}

// operator<< into a file; returns 0 on success
int ref_net_dump(void *net, const char *path) {
  auto *n = static_cast<network::Network *>(net);
  std::ofstream out(path);
  if (!out.is_open()) return 1;
  out << n;
  return 0;
}

// ---- encoder -----------------------------------------------------------------------------------------
// Parses a training-case file with the reference's own loader and returns Piece::code() per square and the target.
int ref_load_case(const char *path, double *codes_out, double *target_out) {
  auto tc = network::load_training_case(path, false);
  std::vector<piece::Piece *> pcs = tc.first->pieces();
  for (int i = 0; i < 64; ++i) codes_out[i] = pcs[i]->code();
  *target_out = tc.second;
  delete tc.first;
  return 0;
}

// code() of a board built from int8 codes through the text format (round trip of the encoder)
int ref_codes_to_inputs(const int8_t *codes, double *inputs_out) {
  game::Board *b = board_from_codes(codes);
  if (b == nullptr) return 1;
  std::vector<piece::Piece *> pcs = b->pieces();
  for (int i = 0; i < 64; ++i) inputs_out[i] = pcs[i]->code();
  delete b;
  return 0;
}

// ---- forward -----------------------------------------------------------------------------------------
int ref_predict_codes(void *net, const int8_t *codes, long n_boards, double *out) {
  auto *n = static_cast<network::Network *>(net);
  for (long b = 0; b < n_boards; ++b) {
    game::Board *board = board_from_codes(codes + 64 * b);
    if (board == nullptr) return 1;
    out[b] = n->predictPosition(board);
    delete board;
  }
  return 0;
}

// Timed forward over pre-built boards, `threads` host threads, one Network::clone() per thread over disjoint
// shards (predictPosition is re-entrant, network.cpp:93-94). Returns seconds of wall time for the predict loop only.
double ref_time_predict(void *net, const int8_t *codes, long n_boards, int threads, int reps, double *out) {
  auto *n = static_cast<network::Network *>(net);
  std::vector<game::Board *> boards(n_boards);
  for (long b = 0; b < n_boards; ++b) boards[b] = board_from_codes(codes + 64 * b);
  if (threads < 1) threads = 1;
  std::vector<network::Network *> nets(threads);
  for (int t = 0; t < threads; ++t) nets[t] = n->clone();
  std::atomic<int> ready(0);
  std::atomic<bool> go(false);
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&, t] {
      long lo = n_boards * t / threads, hi = n_boards * (t + 1) / threads;
      ready.fetch_add(1);
      while (!go.load()) {}
      for (int r = 0; r < reps; ++r)
        for (long b = lo; b < hi; ++b) out[b] = nets[t]->predictPosition(boards[b]);
    });
  while (ready.load() < threads) {}
  auto t0 = std::chrono::steady_clock::now();
  go.store(true);
  for (auto &th : pool) th.join();
  auto t1 = std::chrono::steady_clock::now();
  for (auto *b : boards) delete b;
  for (auto *c : nets) delete c;
  return std::chrono::duration<double>(t1 - t0).count();
}

// ---- training ----------------------------------------------------------------------------------------
int ref_optimize_codes(void *net, const int8_t *codes, double target, double *lambda, double max_const, double rate) {
  auto *n = static_cast<network::Network *>(net);
  game::Board *board = board_from_codes(codes);
  if (board == nullptr) return 1;
  network::Optimizer::optimize(n, board, target, *lambda, max_const, rate);
  delete board;
  return 0;
}

// train.cpp:39-68 restated around the reference's own optimize/dropout for a fixed case order:
// lambda reset + dropout when lambda leaves [1e-4, 10]. Returns the number of resets.
int ref_train_sequence(void *net, const int8_t *codes, const double *targets, long n_cases, double *lambda) {
  auto *n = static_cast<network::Network *>(net);
  const double LAMBDA_MIN = 0.0001, LAMBDA_MAX = 10.0;
  int resets = 0;
  for (long c = 0; c < n_cases; ++c) {
    game::Board *board = board_from_codes(codes + 64 * c);
    if (board == nullptr) return -1;
    network::Optimizer::optimize(n, board, targets[c], *lambda);
    delete board;
    if (*lambda < LAMBDA_MIN || LAMBDA_MAX < *lambda) {
      *lambda = network::Optimizer::DEFAULT_INITIAL_LAMBDA;
      network::Optimizer::dropout(n);
      ++resets;
    }
  }
  return resets;
}

double ref_time_train(void *net, const int8_t *codes, const double *targets, long n_cases, double *lambda) {
  auto *n = static_cast<network::Network *>(net);
  std::vector<game::Board *> boards(n_cases);
  for (long b = 0; b < n_cases; ++b) boards[b] = board_from_codes(codes + 64 * b);
  const double LAMBDA_MIN = 0.0001, LAMBDA_MAX = 10.0;
  auto t0 = std::chrono::steady_clock::now();
  for (long c = 0; c < n_cases; ++c) {
    network::Optimizer::optimize(n, boards[c], targets[c], *lambda);
    if (*lambda < LAMBDA_MIN || LAMBDA_MAX < *lambda) {
      *lambda = network::Optimizer::DEFAULT_INITIAL_LAMBDA;
      network::Optimizer::dropout(n);
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  for (auto *b : boards) delete b;
  return std::chrono::duration<double>(t1 - t0).count();
}

void ref_dropout(void *net, double rate) { network::Optimizer::dropout(static_cast<network::Network *>(net), rate); }

// math::random(): the default-seeded generator owned by math_util.cpp (math_util.h:35-36 makes it per-TU static,
// so init::initializeRNG never reaches it)
double ref_math_random() { return math::random(); }

// ---- Decider::prediction on a position given as a board text file ---------------------------------------
// Returns eval, number of moves and (up to cap) logits in std::map<Move,double> iteration order.
int ref_prediction_file(void *net, const char *board_path, int white_to_move, double *eval_out, double *logits_out,
                        int cap) {
  auto *n = static_cast<network::Network *>(net);
  auto *g = new game::Game(8, 8);
  g->board()->loadFromFile(board_path);
  if (!white_to_move) g->_current_player_color = piece::PieceColor::BLACK;
  g->generatePossibleMoveVectors();  // refresh cached move lists for the loaded position (game.cpp:892-897)
  auto pr = n->prediction(g);
  *eval_out = pr.first;
  int k = 0;
  for (auto &kv : pr.second) {
    if (k < cap) logits_out[k] = kv.second;
    ++k;
  }
  delete g;
  return k;
}

}  // extern "C"

// ---- rules engine: recorded random games (golden data for the fast move generator) ----------
// Plays one game with the reference's own Game / Board / Move code from the given start-position file. At every ply
// records the board as codes, the side to move, the full legal move list in the reference's generation order
// (game.cpp:228-263: from-square ascending, to-square ascending, promotions Q,R,N,B) and the move played. Moves are
// chosen at random with a bias towards castling, en passant, double steps and promotions so the rare rules are covered.
// Layout of `moves_out`: per move 5 ints (r1, c1, r2, c2, promotion code: 0 none, 3 Q, 4 R, 6 N, 7 B).
// Returns the number of plies recorded; *result_out = GameResult::evaluate(), *over_out = isOver().
extern "C" int ref_record_game(const char *start_path, unsigned seed, int max_plies, int8_t *codes_out, int8_t *side_out,
                               int *n_moves_out, int *moves_out, int moves_cap, int *chosen_out, int *nocap_out,
                               int *result_out, int *over_out) {
  auto *blank = new game::Game(8, 8);
  blank->board()->loadFromFile(start_path);
  game::Game *g = blank->clone();
  std::mt19937 rng(seed);
  int ply = 0, mpos = 0;
  auto promo_code = [](piece::PieceType t) -> int {
    if (t.isQueen()) return 3;
    if (t.isRook()) return 4;
    if (t.isKnight()) return 6;
    if (t.isBishop()) return 7;
    return 0;
  };
  while (ply < max_plies && !g->isOver()) {
    std::vector<piece::Piece *> pcs = g->board()->pieces();
    for (int i = 0; i < 64; ++i) codes_out[ply * 64 + i] = (int8_t) std::lround(pcs[i]->code() * 2.5);
    side_out[ply] = (int8_t) g->getCurrentColor().value();
    nocap_out[ply] = g->_moves_since_last_capture;
    std::vector<game::Move> moves = g->possibleMoves(g->getCurrentColor());
    n_moves_out[ply] = (int) moves.size();
    if (moves.empty() || mpos + (int) moves.size() > moves_cap) break;
    std::vector<int> special;
    for (size_t k = 0; k < moves.size(); ++k) {
      const game::Move &m = moves[k];
      int *o = moves_out + 5 * (mpos + (int) k);
      o[0] = m.startingRow(); o[1] = m.startingColumn(); o[2] = m.endingRow(); o[3] = m.endingColumn();
      o[4] = promo_code(m.pawn_promotion_type());
      piece::Piece *p = g->board()->getPiece(o[0], o[1]);
      const bool pawn = p->type().isPawn();
      const bool castle = p->type().isKing() && std::abs(o[1] - o[3]) == 2;
      const bool ep = pawn && o[1] != o[3] && g->board()->getPiece(o[2], o[3])->type().isEmpty();
      const bool dbl = pawn && std::abs(o[0] - o[2]) == 2;
      if (castle || ep || o[4] != 0 || (dbl && rng() % 3 == 0)) special.push_back((int) k);
    }
    int pick;
    if (!special.empty() && rng() % 2 == 0) pick = special[rng() % special.size()];
    else pick = (int) (rng() % moves.size());
    chosen_out[ply] = pick;
    mpos += (int) moves.size();
    if (!g->tryMove(moves[pick])) break;
    ++ply;
  }
  // final position (after the last move) so the replay can check it too
  std::vector<piece::Piece *> pcs = g->board()->pieces();
  for (int i = 0; i < 64; ++i) codes_out[ply * 64 + i] = (int8_t) std::lround(pcs[i]->code() * 2.5);
  side_out[ply] = (int8_t) g->getCurrentColor().value();
  nocap_out[ply] = g->_moves_since_last_capture;
  n_moves_out[ply] = (int) g->possibleMoves(g->getCurrentColor()).size();
  *result_out = g->getResult().evaluate();
  *over_out = g->isOver() ? 1 : 0;
  return ply;
}

// perft on the reference engine (legal move generation + tryMove), for cross-checking node counts
static long ref_perft_rec(game::Game *g, int depth) {
  if (depth == 0) return 1;
  if (g->isOver()) return 0;
  long n = 0;
  for (auto &m : g->possibleMoves(g->getCurrentColor())) {
    game::Game *c = g->clone();
    if (c->tryMove(m)) n += ref_perft_rec(c, depth - 1);
    delete c;
  }
  return n;
}
extern "C" long ref_perft(const char *start_path, int depth) {
  auto *blank = new game::Game(8, 8);
  blank->board()->loadFromFile(start_path);
  game::Game *g = blank->clone();
  return ref_perft_rec(g, depth);
}

// ---- training-case rule ---------------------------------------------------------------
double get_overall_result(double mcts_result, double game_result, double scale);  // make_cases.cpp:36 (external linkage)

// The reference's two-stage target: SIMULATION_BOARD_SAVE_PROCEDURE (initialization.cpp:221-227, ratio 1.0 so the
// board is always kept) maps the searched node's value to [-1,1]; get_overall_result mixes in the game result.
extern "C" int ref_case_target(double node_value, int game_result, double *bounded_out, double *target_out) {
  settings::PRINT_INITIALIZATION_DEBUG_INFORMATION = false;
  init::updateTrainingParameters([] { return true; }, 20, 1.0);
  std::vector<std::pair<game::Board *, double>> kept;
  const int8_t empty[64] = {0};
  game::Board *b = board_from_codes(empty);   // a parsed board: a bare Board(8,8) holds no piece objects to clone
  settings::SIMULATION_BOARD_SAVE_PROCEDURE(kept, b, node_value);
  delete b;
  if (kept.size() != 1) return 1;
  *bounded_out = kept[0].second;
  *target_out = get_overall_result(kept[0].second, (double) game_result, 10.0);
  delete kept[0].first;
  return 0;
}

// Writes one training case with the reference's writer (Board::saveToFile + the target line of
// network::save_training_case, make_cases.cpp:113-121), at an explicit path.
extern "C" int ref_save_case(const int8_t *codes, double target, const char *path) {
  game::Board *b = board_from_codes(codes);
  if (b == nullptr) return 1;
  b->saveToFile(path, [&](std::ostream &out) -> void { out << string::from_double(target) << std::endl; }, false);
  delete b;
  return 0;
}

// Decider::prediction (decider.cpp:31-61) on a position given as codes: the evaluation, and per legal move (in the
// std::map's order) its from / to squares and its logit.
extern "C" int ref_prediction_codes(void *net, const int8_t *codes, int white_to_move, double *eval_out, int *from_out, int *to_out,
                                    double *logits_out, int cap) {
  auto *n = static_cast<network::Network *>(net);
  auto *g = new game::Game(8, 8);
  std::stringstream ss;
  ss << "8 8\n";
  for (int i = 0; i < 64; ++i) {
    const char *txt = piece_text(codes[i]);
    if (txt == nullptr) { delete g; return -1; }
    ss << i << " - " << txt << "\n";
  }
  game::Board *gb = g->board();
  ss >> gb;
  if (!white_to_move) g->_current_player_color = piece::PieceColor::BLACK;
  g->generatePossibleMoveVectors();
  auto pr = n->prediction(g);
  *eval_out = pr.first;
  int k = 0;
  for (auto &kv : pr.second) {
    if (k < cap) {
      from_out[k] = kv.first.startingRow() * 8 + kv.first.startingColumn();
      to_out[k] = kv.first.endingRow() * 8 + kv.first.endingColumn();
      logits_out[k] = kv.second;
    }
    ++k;
  }
  delete g;
  return k;
}

// ---- tree::MCTS on a position given as codes (golden data for the GPU-resident search, cab_mcts.cu) ---------------
// The reference's own sequential search: root = new Node(colour); expand_node(root) (NO Dirichlet noise — the noise is
// added by run_mcts_multithreaded, tree.cpp:182-183, from an un-vendored GSL); MCTS::mcts() with `sims` playouts of
// SIMULATION_SEARCH_DEPTH plies (tree.cpp:202-262). Per root child, in std::map<Move> order: from / to square, prior,
// visit count, value sum; the root's own visit count and value sum. Returns the number of children.
extern "C" int ref_mcts_root(void *net, const int8_t *codes, int white_to_move, int no_capture, int sims, int cap, int *from_out,
                             int *to_out, double *prior_out, int *visits_out, double *value_sum_out, int *root_visits, double *root_value_sum) {
  auto *n = static_cast<network::Network *>(net);
  auto *g = new game::Game(8, 8);
  std::stringstream ss;
  ss << "8 8\n";
  for (int i = 0; i < 64; ++i) {
    const char *txt = piece_text(codes[i]);
    if (txt == nullptr) { delete g; return -1; }
    ss << i << " - " << txt << "\n";
  }
  game::Board *gb = g->board();
  ss >> gb;
  if (!white_to_move) {
    g->_current_player_color = piece::PieceColor::BLACK;
    gb->_move_count.store(1);   // Board::doMove derives the next colour from the move count's parity (game.cpp:271-276)
  }
  g->_moves_since_last_capture = no_capture;
  g->generatePossibleMoveVectors();
  auto *root = new tree::Node(g->getCurrentColor());
  tree::MCTS::expand_node(root, g, n);
  std::atomic_int iterations{sims}, finished{0};
  tree::MCTS::mcts(g, n, root, iterations, finished);
  int k = 0;
  for (auto &kv : root->_children) {
    if (k < cap) {
      from_out[k] = kv.first.startingRow() * 8 + kv.first.startingColumn();
      to_out[k] = kv.first.endingRow() * 8 + kv.first.endingColumn();
      prior_out[k] = kv.second->_priority;
      visits_out[k] = kv.second->_visit_count;
      value_sum_out[k] = kv.second->_value_sum;
    }
    ++k;
  }
  *root_visits = root->_visit_count;
  *root_value_sum = root->_value_sum;
  delete root;
  delete g;
  return k;
}

// The same search with every playout written down (debugging aid for the GPU-resident search): the loop of
// tree::MCTS::mcts (tree.cpp:202-262) restated around the reference's own select_optimal_move / expand_node / tryMove /
// propagate_result. Per playout: trace[0] = plies walked, trace[1 + 2i] = from, trace[2 + 2i] = to, values_out = leaf value,
// over_out = isOver at the leaf. Stride of `trace` is 20 ints.
extern "C" int ref_mcts_trace(void *net, const int8_t *codes, int white_to_move, int no_capture, int sims, int *trace, double *values_out,
                              int *over_out) {
  auto *n = static_cast<network::Network *>(net);
  auto *g = new game::Game(8, 8);
  std::stringstream ss;
  ss << "8 8\n";
  for (int i = 0; i < 64; ++i) ss << i << " - " << piece_text(codes[i]) << "\n";
  game::Board *gb = g->board();
  ss >> gb;
  if (!white_to_move) {
    g->_current_player_color = piece::PieceColor::BLACK;
    gb->_move_count.store(1);   // Board::doMove derives the next colour from the move count's parity (game.cpp:271-276)
  }
  g->_moves_since_last_capture = no_capture;
  g->generatePossibleMoveVectors();
  auto *root = new tree::Node(g->getCurrentColor());
  tree::MCTS::expand_node(root, g, n);
  for (int s = 0; s < sims; ++s) {
    tree::Node *node = root;
    game::Game *clone = g->clone();
    std::vector<tree::Node *> path = {root};
    int *t = trace + 20 * s;
    t[0] = 0;
    for (int i = 0; i < tree::MCTS::SIMULATION_SEARCH_DEPTH; ++i) {
      if (clone->isOver()) break;
      if (!node->expanded()) tree::MCTS::expand_node(node, clone, n);
      auto optimal = tree::MCTS::select_optimal_move(node);
      node = optimal.second;
      t[1 + 2 * t[0]] = optimal.first.startingRow() * 8 + optimal.first.startingColumn();
      t[2 + 2 * t[0]] = optimal.first.endingRow() * 8 + optimal.first.endingColumn();
      ++t[0];
      if (!clone->tryMove(optimal.first)) break;
      path.push_back(node);
    }
    const double cc = clone->getCurrentColor().value();
    double value;
    if (clone->isOver()) value = (cc * clone->getResult().evaluate() + 1.0) / 2.0;
    else {
      value = clone->board()->score([&](piece::Piece *piece) -> double { return cc * piece->color().value() * piece->type().minimaxValue(); });
      value = 1.0 / (1.0 + exp(-value / 27.18));
    }
    values_out[s] = value;
    over_out[s] = clone->isOver() ? 1 : 0;
    tree::MCTS::propagate_result(path, value, clone->getCurrentColor());
    delete clone;
  }
  delete root;
  delete g;
  return 0;
}
