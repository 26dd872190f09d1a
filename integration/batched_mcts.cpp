// batched_mcts.cpp — MCTS with virtual-loss leaf batching over the REFERENCE's own rules engine.
//
// Search semantics restated from /root/reference/src/mcts_network/tree.cpp (not copied; cited per item):
//   playout            :202-262   depth-limited descent (SIMULATION_SEARCH_DEPTH plies), every unexpanded node met on
//                                 the way is expanded, leaf value = game result or sigmoid(material / 27.18)
//   PUCT               :276-283   (ln((N+19652+1)/19652)+1.25) * sqrt(N)/(n+1) * prior + value
//   expansion / priors :294-306   prior_i = exp(cm*logit_i)/sum, children in std::map<Move> order (:86-95)
//   back-propagation   :264-269   visits += 1, value_sum += |v or 1-v| by colour
//   move logits        decider.cpp:43-58, including its inverted branch (bug-compatible by default)
// What is new: a playout that reaches an unexpanded node does not evaluate it on the spot. It puts a VIRTUAL LOSS on
// its path (visits += 1 with no value, so sibling playouts steer elsewhere), pushes the boards it needs into the
// leaf queue (cab_leafq_push) and parks; the scheduler keeps launching playouts over all games until `in_flight`
// are parked, then ONE cab_leafq_flush evaluates every queued board on the GPU, the parked playouts are expanded
// (priors through K7) and resumed. With in_flight == 1 the schedule degenerates to the reference's sequential loop,
// which is what the parity mode checks against the reference's own tree::MCTS::mcts().
//
// Built only where /root/reference exists (integration/Makefile), with -fno-access-control so the parity mode can
// call the reference's private MCTS::mcts / expand_node. Prints one JSON object.
//
//   batched_mcts <latest.txt> <start_position.txt> <games> <playouts_per_game> <in_flight> <mode> [network_priors]
//     mode "parity": 1 game, compares root visit counts with the reference's sequential mcts()
//     mode "bench" : throughput / batching statistics
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "chess/game.h"
#include "chessai_b200.h"
#include "mcts_network/decider.h"
#include "mcts_network/network.h"
#include "mcts_network/tree.h"

namespace {

struct BNode {
  double prior = 0.0;
  int visits = 0;          // includes virtual losses while playouts are parked below
  int virtual_loss = 0;
  double value_sum = 0.0;
  bool expanded = false;
  double color = 1.0;      // colour to play at this node (+1 white, -1 black)
  std::vector<std::pair<game::Move, std::unique_ptr<BNode>>> children;  // std::map<Move> iteration order
  double value() const { const int n = visits - virtual_loss; return n > 0 ? value_sum / n : 0.0; }   // tree.cpp:74-76
};

double puct(const BNode &parent, const BNode &child) {                                                  // tree.cpp:276-283
  double base = std::log((parent.visits + 19652.0 + 1.0) / 19652.0) + 1.25;
  base *= std::sqrt((double) parent.visits) / (child.visits + 1);
  return base * child.prior + child.value();
}

void encode(game::Board *b, int8_t *codes) {                                                            // network.cpp:119-123
  std::vector<piece::Piece *> pcs = b->pieces();
  for (int i = 0; i < 64; ++i) codes[i] = (int8_t) std::lround(pcs[i]->code() * 2.5);
}

// one expansion request: the moves of the node and which queue slots carry their boards
struct Pending {
  std::vector<game::Move> moves;
  std::vector<long> slot;          // per move: queue slot of the child board, or -1 when the logit is the constant
  long parent_slot = -1;
};

struct Playout {
  game::Game *clone = nullptr;
  BNode *node = nullptr;
  std::vector<BNode *> path;
  int depth = 0;
  Pending pend;
  bool parked = false;
};

struct Searcher {
  cab_leafq *q = nullptr;
  int search_depth = 8;
  bool network_priors = false;   // false: decider.cpp:52-57 as shipped (logit = cm*20 when the opponent can reply)
  long evals = 0, expansions = 0;

  // Decider::prediction's rules part (decider.cpp:39-58): collect the boards whose evaluation is needed
  void request_expansion(Playout &p) {
    game::Game *g = p.clone;
    game::Board *board = g->board();
    piece::PieceColor col = g->getCurrentColor();
    int8_t codes[64];
    encode(board, codes);
    cab_leafq_push(q, codes, 1, &p.pend.parent_slot);
    p.pend.moves.clear();
    p.pend.slot.clear();
    // std::map collapses the four promotions of a pawn move into the first key (the queen promotion, Q8) while
    // actionMap[move] = ... keeps the value of the LAST one assigned (the bishop promotion), decider.cpp:52-56
    std::map<game::Move, game::Move> ordered;
    std::vector<game::Move> moves = g->possibleMoves(col), reply;
    for (auto &m : moves) {
      if (!m.verify(board)) continue;
      auto it = ordered.find(m);
      if (it == ordered.end()) ordered.emplace(m, m);
      else it->second = m;
    }
    for (auto &kv : ordered) {
      const game::Move &m = kv.first;
      board->doMove(new game::Move(kv.second), g);      // the board whose evaluation becomes the key's logit
      if (col.isWhite()) board->getPossibleMoves(nullptr, &reply);
      else board->getPossibleMoves(&reply, nullptr);
      long slot = -1;
      if (reply.empty() || network_priors) {
        encode(board, codes);
        cab_leafq_push(q, codes, 1, &slot);
      }
      reply.clear();
      board->undoMove(g);
      p.pend.moves.push_back(m);
      p.pend.slot.push_back(slot);
    }
    ++expansions;
  }

  // after the flush: logits -> priors (K7 math, tree.cpp:299-306) -> children
  void finish_expansion(Playout &p, const double *results) {
    BNode *n = p.node;
    const double cm = n->color;
    const size_t k = p.pend.moves.size();
    std::vector<double> w(k);
    double sum = 0.0;
    for (size_t i = 0; i < k; ++i) {
      const double logit = p.pend.slot[i] >= 0 ? cm * results[p.pend.slot[i]] : cm * 20.0;   // decider.cpp:52-57
      w[i] = std::exp(cm * logit);
      sum += w[i];
    }
    n->children.clear();
    for (size_t i = 0; i < k; ++i) {
      auto c = std::make_unique<BNode>();
      c->prior = w[i] / sum;
      c->color = -cm;
      n->children.emplace_back(p.pend.moves[i], std::move(c));
    }
    n->expanded = true;
    evals += 1 + (long) k;
  }

  // advance a playout until it parks (needs an expansion) or finishes; returns true when finished
  bool advance(Playout &p) {
    while (p.depth < search_depth) {
      if (p.clone->isOver()) break;
      if (!p.node->expanded) {
        request_expansion(p);
        for (BNode *n : p.path) { n->visits += 1; n->virtual_loss += 1; }   // virtual loss on the whole path
        p.parked = true;
        return false;
      }
      BNode *best = nullptr;
      const game::Move *best_move = nullptr;
      double max_score = -1.0;                                              // Node::selectOptimalMove, tree.cpp:103-118
      for (auto &c : p.node->children) {
        const double s = puct(*p.node, *c.second);
        if (s > max_score) { max_score = s; best = c.second.get(); best_move = &c.first; }
      }
      if (best == nullptr || !p.clone->tryMove(*best_move)) break;
      p.node = best;
      p.path.push_back(best);
      ++p.depth;
    }
    // leaf value (tree.cpp:231-253)
    const double cc = p.clone->getCurrentColor().value();
    double value;
    if (p.clone->isOver()) value = (cc * p.clone->getResult().evaluate() + 1.0) / 2.0;
    else {
      value = p.clone->board()->score([&](piece::Piece *pc) -> double { return cc * pc->color().value() * pc->type().minimaxValue(); });
      value = 1.0 / (1.0 + std::exp(-value / 27.18));
    }
    for (BNode *n : p.path) {                                               // propagate_result, tree.cpp:264-269
      n->visits += 1;
      n->value_sum += std::fabs(cc == n->color ? value : 1.0 - value);
    }
    delete p.clone;
    p.clone = nullptr;
    return true;
  }

  void unpark(Playout &p, const double *results) {
    for (BNode *n : p.path) { n->visits -= 1; n->virtual_loss -= 1; }
    finish_expansion(p, results);
    p.parked = false;
  }
};

}  // namespace

int main(int argc, char **argv) {
  if (argc < 7) { std::fprintf(stderr, "usage: %s latest.txt start.txt games playouts in_flight parity|bench [network_priors]\n", argv[0]); return 2; }
  const int games = std::atoi(argv[3]), playouts = std::atoi(argv[4]), in_flight = std::atoi(argv[5]);
  const std::string mode = argv[6];
  const bool network_priors = argc > 7 && std::atoi(argv[7]) != 0;

  cab_net *net = nullptr;
  if (cab_net_load_text(argv[1], 0, &net) != CAB_OK) { std::fprintf(stderr, "%s\n", cab_last_error()); return 1; }
  Searcher S;
  S.network_priors = network_priors;
  S.search_depth = tree::MCTS::SIMULATION_SEARCH_DEPTH;
  if (cab_leafq_create(net, 1 << 18, 4096, 4, &S.q) != CAB_OK) { std::fprintf(stderr, "%s\n", cab_last_error()); return 1; }

  auto *blank = new game::Game(8, 8);
  blank->board()->loadFromFile(argv[2]);
  std::vector<game::Game *> roots_game(games);
  std::vector<std::unique_ptr<BNode>> roots(games);
  for (int g = 0; g < games; ++g) {
    roots_game[g] = blank->clone();
    roots[g] = std::make_unique<BNode>();
    roots[g]->color = roots_game[g]->getCurrentColor().value();
  }

  // ---- the batched search: round-robin over games, park on expansion, flush when in_flight are parked ----
  std::vector<int> launched(games, 0), done(games, 0);
  std::vector<Playout> parked;
  long flushes = 0, max_batch = 0, total_batched = 0;
  auto t0 = std::chrono::steady_clock::now();
  bool work_left = true;
  while (work_left) {
    work_left = false;
    for (int g = 0; g < games && (int) parked.size() < in_flight; ++g) {
      while (launched[g] < playouts && (int) parked.size() < in_flight) {
        // at most one parked playout per unexpanded root: start the next playout of a game only when its root exists
        if (!roots[g]->expanded && launched[g] > done[g]) break;
        Playout p;
        p.clone = roots_game[g]->clone();
        p.node = roots[g].get();
        p.path = {p.node};
        ++launched[g];
        if (S.advance(p)) ++done[g];
        else parked.push_back(std::move(p));
      }
    }
    if (!parked.empty()) {
      long n_eval = 0;
      if (cab_leafq_flush(S.q, &n_eval) != CAB_OK) { std::fprintf(stderr, "%s\n", cab_last_error()); return 1; }
      const double *res = nullptr;
      cab_leafq_results(S.q, &res);
      ++flushes;
      total_batched += n_eval;
      max_batch = std::max(max_batch, n_eval);
      std::vector<Playout> still;
      // two playouts may have parked on the SAME node (no virtual loss can separate them before the first
      // expansion of a root): expand once, resume both
      for (auto &p : parked) {
        for (BNode *n : p.path) { n->visits -= 1; n->virtual_loss -= 1; }
        if (!p.node->expanded) S.finish_expansion(p, res);
        p.parked = false;
      }
      for (auto &p : parked) {
        // find the game this playout belongs to through its root
        int g = 0;
        for (; g < games; ++g) if (roots[g].get() == p.path[0]) break;
        if (S.advance(p)) ++done[g];
        else still.push_back(std::move(p));
      }
      parked.swap(still);
    }
    for (int g = 0; g < games; ++g) if (done[g] < playouts) work_left = true;
  }
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  // ---- invariants ----
  bool ok = true;
  for (int g = 0; g < games; ++g) {
    int child_visits = 0;
    double prior_sum = 0.0;
    for (auto &c : roots[g]->children) { child_visits += c.second->visits; prior_sum += c.second->prior; if (c.second->virtual_loss != 0) ok = false; }
    if (roots[g]->visits != playouts || roots[g]->virtual_loss != 0) ok = false;
    if (std::fabs(prior_sum - 1.0) > 1e-12) ok = false;
    if (child_visits > playouts) ok = false;
  }

  // ---- parity with the reference's sequential search (mode parity: 1 game, in_flight 1, no root noise) ----
  int parity_checked = 0, parity_equal = 0;
  if (mode == "parity") {
    network::Network *ref_net = network::Network::loadFromFile(argv[1]);   // integration/network_b200.cpp on the same GPU
    game::Game *g = blank->clone();
    auto *root = new tree::Node(g->getCurrentColor());
    tree::MCTS::expand_node(root, g, ref_net);
    std::atomic_int iters(playouts), finished(0);
    tree::MCTS::mcts(g, ref_net, root, iters, finished);                   // the reference's own playout loop
    parity_checked = 1;
    parity_equal = (int) root->_children.size() == (int) roots[0]->children.size() && root->_visit_count == roots[0]->visits;
    size_t i = 0;
    for (auto &kv : root->_children) {
      if (i >= roots[0]->children.size()) { parity_equal = 0; break; }
      const BNode &mine = *roots[0]->children[i].second;
      if (!(kv.first == roots[0]->children[i].first) || kv.second->_visit_count != mine.visits ||
          std::fabs(kv.second->_value_sum - mine.value_sum) > 1e-9 || std::fabs(kv.second->_priority - mine.prior) > 1e-12)
        parity_equal = 0;
      ++i;
    }
    delete root;
    delete ref_net;
  }

  long launches = 0, q_flushes = 0, evaluated = 0;
  cab_leafq_stats(S.q, &launches, &q_flushes, &evaluated);
  std::printf("{\"games\": %d, \"playouts_per_game\": %d, \"in_flight\": %d, \"network_priors\": %d, \"invariants_ok\": %d, "
              "\"flushes\": %ld, \"boards_evaluated\": %ld, \"avg_boards_per_flush\": %.1f, \"max_boards_per_flush\": %ld, "
              "\"gpu_launches\": %ld, \"expansions\": %ld, \"seconds\": %.3f, \"playouts_per_s\": %.1f, \"boards_per_s\": %.1f, "
              "\"parity_checked\": %d, \"parity_equal\": %d}\n",
              games, playouts, in_flight, (int) network_priors, (int) ok, flushes, evaluated,
              flushes ? (double) total_batched / flushes : 0.0, max_batch, launches, S.expansions, secs,
              (double) games * playouts / secs, evaluated / secs, parity_checked, parity_equal);
  cab_leafq_destroy(S.q);
  cab_net_destroy(net);
  return ok ? 0 : 3;
}
