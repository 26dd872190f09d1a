// selfplay.cpp — MCTS self-play with virtual-loss leaf batching on the B200 evaluator (BASELINE.json configs[2]).
//
// Many concurrent games per GPU; every host thread owns a slice of the games, its own trees and its own leaf queue
// (cab_leafq_*), so threads never synchronise with each other and their GPU launches overlap. Per thread:
//   repeat: start / resume playouts round-robin over its games; a playout that meets an unexpanded node puts a
//           VIRTUAL LOSS on its path, pushes the node's board and all child boards (one contiguous push) and parks;
//           when `in_flight` playouts are parked (or nothing else can start) ONE flush evaluates them all on the GPU,
//           priors are formed, the parked playouts resume.
// Search semantics = the reference's (/root/reference/src/mcts_network/tree.cpp, restated; integration/batched_mcts.cpp
// proves the same scheduler equal to tree::MCTS::mcts() on the reference's rules code): PUCT :276-283, priors
// exp(cm*logit)/sum :299-306, depth-limited playouts that expand on the way :202-262, leaf value = result or
// sigmoid(material/27.18) :231-253, Dirichlet(0.2) noise mixed 25 % into the root priors :318-347, the move played is
// the root child with the best PUCT score :195. Rules: host/fastchess.hpp (the reference's rules, verified move for move).
// Child logits follow decider.cpp:43-58: network_priors = 0 keeps the shipped inverted branch (constant 20 unless the
// opponent has no reply), 1 evaluates every child with the network.
//
// Training cases (host/cases.hpp; make_cases.cpp:36-59, initialization.cpp:221-227): after every move the new board is
// kept with probability save_ratio together with clamp(2*root_value-1); when the game ends each kept board gets the
// target 10*(4*b + result)/5 and is appended to the packed shard `cases_path`.
//
//   selfplay <latest.txt> <games> <playouts_per_move> <moves_per_game> <threads> <in_flight_per_thread> <network_priors>
//            [device] [seed] [cases_path] [save_ratio] [precision: fp64 | f16] [roots]
// (roots = independent trees per move, merged like the reference's root-parallel threads; each gets playouts / roots)
// (cases_path "-" = none; precision f16 = the reduced-precision fused tcgen05 evaluator, see cab_leafq_set_precision)
// Prints one JSON object. No CPU fallback: needs libchessai_b200.so and a GPU.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

#include "../../include/chessai_b200.h"
#include "cases.hpp"
#include "fastchess.hpp"

using namespace fastchess;

namespace {

constexpr int kSearchDepth = 8;  // tree.cpp:145 SIMULATION_SEARCH_DEPTH

struct Node {
  double prior = 0.0, value_sum = 0.0;
  int visits = 0, vloss = 0;
  int first_child = -1;
  uint16_t n_children = 0;
  int8_t color = 1;
  bool expanded = false;
  bool pending = false;            // children allocated, their evaluations are in the queue
  bool eval = false;               // (as a child) its board was sent to the network for the prior logit
  int8_t state = 0;                // 0 = not examined, 1 = the side to move has a legal move, 2 = game over here
  int8_t result = 0;               // winner when state == 2 (0 = draw)
  Move move{0, 0, 0};
  double value() const { const int n = visits - vloss; return n > 0 ? value_sum / n : 0.0; }
};
// The per-child part of the PUCT score, kept apart from the nodes and contiguous per sibling group: the selection loop
// (the hottest loop on the host, ~30 children x 8 levels per playout over trees far larger than the caches) then
// streams 8 bytes per child instead of a whole node.
struct ChildStat {
  float q = 0.0f, u = 0.0f;        // value() and prior / (visits + 1); single precision: 8 bytes per child
};

struct Playout {
  Position pos;
  int tree = 0, node = 0, depth = 0;   // tree = game * roots + root
  int path[kSearchDepth + 2];
  int path_len = 0;
  long slot = -1;     // first queue slot of the expansion this playout requested (parent board, then the evaluated
                      // children); -1 when it only waits for an expansion another playout requested
};

// One search tree. A move is searched by `roots` independent trees from the same position (the reference's
// root-parallel threads, tree.cpp:153-196: DEFAULT_NUM_THREADS = 4 roots of NUM_SIMULATIONS_PER_THREAD playouts each, own
// Dirichlet noise), merged like Node::combineNodes (tree.cpp:122-142) before the move is chosen.
struct Tree {
  std::vector<Node> arena;
  std::vector<ChildStat> stat;     // parallel to arena
  void refresh(int i) { const Node &n = arena[i]; stat[i].q = (float) n.value(); stat[i].u = (float) (n.prior / (n.visits + 1)); }
  int launched = 0, done = 0;
};

struct GameState {
  Position pos;
  int moves_played = 0;
  bool finished = false;
  std::vector<cases::Case> kept;   // boards of this game awaiting the game result
};

struct Stats {
  long playouts = 0, boards = 0, flushes = 0, expansions = 0, moves = 0, max_flush = 0;
  double flush_seconds = 0.0;    // time this thread spent inside cab_leafq_flush (waiting for the GPU)
};

struct Worker {
  cab_net *net;
  cab_leafq *q = nullptr;
  std::vector<GameState> games;
  std::vector<Tree> trees;         // games.size() * roots, the roots of a game adjacent
  int roots = 1;
  int playouts_per_move, moves_per_game, in_flight;
  bool network_priors;
  double save_ratio = 0.0;
  int precision = 0;
  std::vector<cases::Case> finished_cases;
  std::mt19937_64 rng;
  Stats st;

  // tree.cpp:276-283; the parent-only factor is hoisted out of the loop over children
  static double puct_parent(const Node &parent) {
    return (std::log((parent.visits + 19652.0 + 1.0) / 19652.0) + 1.25) * std::sqrt((double) parent.visits);
  }
  static double puct(double parent_factor, const ChildStat &child) { return parent_factor * child.u + child.q; }

  // Game::tryMove without the mate / stalemate scan (game.cpp:847-871): that scan is a full move generation, which
  // the node below repeats anyway when it is expanded; its outcome is cached in Node::state instead.
  static void play_lazy(Position &p, Move mv) {
    const bool capture = apply_move(p.sq, mv);
    p.side = (int8_t) -p.side;
    p.no_capture = capture ? 0 : p.no_capture + 1;
    if (p.no_capture >= 50) { p.over = 1; p.result = 0; }
  }
  static void mark_over(Position &pos, Node &n) {     // Game::updateGameState (game.cpp:872-890) for "no legal move"
    const int ks = king_square(pos.sq, pos.side);
    const bool safe = ks < 0 || !attacked(pos.sq, ks >> 3, ks & 7, pos.side);
    n.state = 2;
    n.result = safe ? 0 : (int8_t) -pos.side;
    pos.over = 1;
    pos.result = n.result;
  }

  // Allocates the node's children and queues their boards. false: the side to move has no legal move.
  bool request_expansion(Playout &p) {
    Move mv[kMaxMoves];
    const int n = generate_moves(p.pos, p.pos.side, mv, /*queen_only=*/true);   // std::map<Move> keeps one promotion (Q8)
    if (n == 0) return false;
    Tree &g = trees[p.tree];
    const int first = (int) g.arena.size();
    g.arena.resize(g.arena.size() + n);
    g.stat.resize(g.arena.size());
    Node &parent = g.arena[p.node];
    parent.first_child = first;
    parent.n_children = (uint16_t) n;
    parent.pending = true;
    static thread_local int8_t buf[(kMaxMoves + 1) * 64];
    std::memcpy(buf, p.pos.sq, 64);
    int n_boards = 1;
    for (int i = 0; i < n; ++i) {
      int8_t *dst = buf + n_boards * 64;
      std::memcpy(dst, p.pos.sq, 64);
      // the LOGIT of a promotion key is the one of the last promotion the reference assigns to it, the bishop's
      // (actionMap[move] = ..., decider.cpp:52-56; cab_prediction_host); the move PLAYED stays the key's queen promotion
      apply_move(dst, mv[i].promo != 0 ? Move{mv[i].from, mv[i].to, 7} : mv[i]);
      bool eval_child = network_priors;
      if (!eval_child) {                       // decider.cpp:52-57: the network is asked only when the opponent has no reply
        Position c;
        std::memcpy(c.sq, dst, 64);
        c.side = (int8_t) -p.pos.side;
        Move tmp[kMaxMoves];
        eval_child = generate_moves(c, c.side, tmp, true) == 0;
      }
      Node &ch = g.arena[first + i];
      ch.color = (int8_t) -parent.color;
      ch.move = mv[i];
      ch.eval = eval_child;
      if (eval_child) ++n_boards;
    }
    if (cab_leafq_push(q, buf, n_boards, &p.slot) != CAB_OK) { std::fprintf(stderr, "%s\n", cab_last_error()); std::exit(1); }
    ++st.expansions;
    return true;
  }

  void finish_expansion(Playout &p, const double *res) {
    if (p.slot < 0) return;                    // this playout only waited for somebody else's expansion
    Tree &g = trees[p.tree];
    Node &n = g.arena[p.node];
    const double cm = n.color;
    const int k = n.n_children, first = n.first_child;
    double sum = 0.0;
    long s = p.slot + 1;
    for (int i = 0; i < k; ++i) {
      Node &c = g.arena[first + i];
      const double logit = c.eval ? cm * res[s++] : cm * 20.0;
      c.prior = std::exp(cm * logit);          // tree.cpp:302
      sum += c.prior;
    }
    for (int i = 0; i < k; ++i) g.arena[first + i].prior /= sum;
    if (p.node == 0 && k > 0) {                // root: Dirichlet(0.2) noise, 25 % (tree.cpp:318-347)
      std::gamma_distribution<double> gamma(0.2, 1.0);
      double th[kMaxMoves], tot = 0.0;
      for (int i = 0; i < k; ++i) { th[i] = gamma(rng); tot += th[i]; }
      for (int i = 0; i < k; ++i) g.arena[first + i].prior = g.arena[first + i].prior * 0.75 + (th[i] / tot) * 0.25;
    }
    for (int i = 0; i < k; ++i) g.refresh(first + i);
    n.pending = false;
    n.expanded = true;
    p.slot = -1;
  }

  // advance until the playout parks (false) or completes (true)
  bool advance(Playout &p) {
    Tree &g = trees[p.tree];
    while (p.depth < kSearchDepth) {
      if (p.pos.over) break;
      Node &n = g.arena[p.node];
      if (n.state == 2) { p.pos.over = 1; p.pos.result = n.result; break; }
      if (!n.expanded) {
        p.slot = -1;
        if (!n.pending) {                      // first playout to get here asks for the expansion, later ones just wait
          if (!request_expansion(p)) { mark_over(p.pos, g.arena[p.node]); break; }
          g.arena[p.node].state = 1;           // (the arena may have moved)
        }
        for (int i = 0; i < p.path_len; ++i) { Node &v = g.arena[p.path[i]]; v.visits += 1; v.vloss += 1; g.refresh(p.path[i]); }
        return false;
      }
      int best = -1;
      double max_score = -1.0;
      const double pf = puct_parent(n);
      for (int i = 0; i < n.n_children; ++i) {
        const double s = puct(pf, g.stat[n.first_child + i]);
        if (s > max_score) { max_score = s; best = n.first_child + i; }
      }
      if (best < 0) break;
      play_lazy(p.pos, g.arena[best].move);
      p.node = best;
      p.path[p.path_len++] = best;
      ++p.depth;
    }
    if (!p.pos.over) {                         // depth limit reached: the leaf-value rule needs to know whether the game ended
      Node &n = g.arena[p.node];
      if (n.state == 2) { p.pos.over = 1; p.pos.result = n.result; }
      else if (n.state == 0) {
        Move tmp[kMaxMoves];
        if (generate_moves(p.pos, p.pos.side, tmp, true) == 0) mark_over(p.pos, n);
        else n.state = 1;
      }
    }
    const double cc = p.pos.side;
    double value;
    if (p.pos.over) value = (cc * p.pos.result + 1.0) / 2.0;
    else value = 1.0 / (1.0 + std::exp(-material(p.pos, p.pos.side) / 27.18));
    for (int i = 0; i < p.path_len; ++i) {
      Node &n = g.arena[p.path[i]];
      n.visits += 1;
      n.value_sum += std::fabs(cc == n.color ? value : 1.0 - value);
      g.refresh(p.path[i]);
    }
    ++st.playouts;
    return true;
  }

  void new_root(Tree &g, const Position &pos) {
    g.arena.clear();
    g.arena.reserve(1 << 14);
    g.arena.emplace_back();
    g.stat.clear();
    g.stat.reserve(1 << 14);
    g.stat.emplace_back();
    g.arena[0].color = pos.side;
    g.launched = g.done = 0;
  }

  void run() {
    if (cab_leafq_create(net, 1 << 18, 4096, 2, &q) != CAB_OK || cab_leafq_set_precision(q, precision) != CAB_OK) {
      std::fprintf(stderr, "%s\n", cab_last_error());
      std::exit(1);
    }
    trees.resize(games.size() * (size_t) roots);
    for (size_t t = 0; t < trees.size(); ++t) new_root(trees[t], games[t / roots].pos);
    const int quota = std::max(1, playouts_per_move / roots);   // playouts per root (NUM_SIMULATIONS_PER_THREAD)
    std::vector<Playout> parked, still;
    parked.reserve(in_flight);
    bool active = true;
    while (active) {
      // launch new playouts round-robin
      bool launched_any = true;
      while (launched_any && (int) parked.size() < in_flight) {
        launched_any = false;
        for (size_t ti = 0; ti < trees.size() && (int) parked.size() < in_flight; ++ti) {
          Tree &g = trees[ti];
          const GameState &game = games[ti / roots];
          if (game.finished || g.launched >= quota) continue;
          if (!g.arena[0].expanded && g.launched > g.done) continue;   // one playout until the root exists
          Playout p;
          p.pos = game.pos;
          p.tree = (int) ti;
          p.node = 0;
          p.path[0] = 0;
          p.path_len = 1;
          ++g.launched;
          launched_any = true;
          if (advance(p)) ++g.done;
          else parked.push_back(p);
        }
      }
      if (!parked.empty()) {
        long n_eval = 0;
        const auto tf = std::chrono::steady_clock::now();
        if (cab_leafq_flush(q, &n_eval) != CAB_OK) { std::fprintf(stderr, "%s\n", cab_last_error()); std::exit(1); }
        st.flush_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - tf).count();
        const double *res = nullptr;
        cab_leafq_results(q, &res);
        ++st.flushes;
        st.boards += n_eval;
        st.max_flush = std::max(st.max_flush, n_eval);
        for (auto &p : parked) {
          Tree &g = trees[p.tree];
          for (int i = 0; i < p.path_len; ++i) { Node &v = g.arena[p.path[i]]; v.visits -= 1; v.vloss -= 1; g.refresh(p.path[i]); }
          finish_expansion(p, res);
        }
        still.clear();
        for (auto &p : parked) {
          if (advance(p)) ++trees[p.tree].done;
          else still.push_back(p);
        }
        parked.swap(still);
      }
      // games whose search is complete: play the move with the best PUCT score (tree.cpp:195) and start a new tree
      active = false;
      for (size_t gi = 0; gi < games.size(); ++gi) {
        GameState &g = games[gi];
        if (g.finished) continue;
        Tree *my = &trees[gi * (size_t) roots];
        bool all_done = true;
        for (int r = 0; r < roots; ++r) all_done = all_done && my[r].done >= quota;
        if (all_done) {
          // Node::combineNodes (tree.cpp:122-142): the roots saw the same position, so their children are the same moves
          // in the same order; priors are averaged, visit counts and value sums added, for the children and the root
          Node root;
          const int k = my[0].arena[0].n_children;
          int best = -1;
          double max_score = -1.0;
          for (int r = 0; r < roots; ++r) { root.visits += my[r].arena[0].visits; root.value_sum += my[r].arena[0].value_sum; }
          const double pf = puct_parent(root);
          for (int i = 0; i < k; ++i) {
            Node c;
            for (int r = 0; r < roots; ++r) {
              const Node &rc = my[r].arena[my[r].arena[0].first_child + i];
              c.prior += rc.prior / roots;
              c.visits += rc.visits;
              c.value_sum += rc.value_sum;
            }
            const double sc = pf * (c.prior / (c.visits + 1)) + c.value();
            if (sc > max_score) { max_score = sc; best = i; }
          }
          if (best >= 0) play(g.pos, my[0].arena[my[0].arena[0].first_child + best].move);
          ++g.moves_played;
          ++st.moves;
          if (save_ratio > 0.0 && std::generate_canonical<double, 53>(rng) < save_ratio) {   // NetworkStorage::saveBoard
            cases::Case c;
            std::memcpy(c.codes, g.pos.sq, 64);
            c.target = cases::bounded_mcts_result(root.value());
            g.kept.push_back(c);
          }
          if (best < 0 || g.pos.over || g.moves_played >= moves_per_game) {
            g.finished = true;
            const double result = g.pos.over ? (double) g.pos.result : 0.0;   // GameResult::evaluate, game.fwd.h:82-95
            for (auto &c : g.kept) { c.target = cases::overall_result(c.target, result); finished_cases.push_back(c); }
            g.kept.clear();
          } else {
            for (int r = 0; r < roots; ++r) new_root(my[r], g.pos);
          }
        }
        if (!g.finished) active = true;
      }
    }
    cab_leafq_destroy(q);
  }
};

}  // namespace

int main(int argc, char **argv) {
  if (argc < 8) {
    std::fprintf(stderr, "usage: %s latest.txt games playouts_per_move moves_per_game threads in_flight_per_thread network_priors [device] [seed]\n", argv[0]);
    return 2;
  }
  const int games = std::atoi(argv[2]), playouts = std::atoi(argv[3]), moves = std::atoi(argv[4]);
  const int threads = std::max(1, std::atoi(argv[5])), in_flight = std::atoi(argv[6]);
  const bool network_priors = std::atoi(argv[7]) != 0;
  const int device = argc > 8 ? std::atoi(argv[8]) : 0;
  const unsigned long long seed = argc > 9 ? std::strtoull(argv[9], nullptr, 10) : 1234;
  const char *cases_path = argc > 10 && std::strcmp(argv[10], "-") != 0 ? argv[10] : nullptr;
  const int precision = argc > 12 && std::strcmp(argv[12], "f16") == 0 ? 1 : 0;
  const int roots = argc > 13 ? std::max(1, std::atoi(argv[13])) : 1;
  const double save_ratio = cases_path ? (argc > 11 ? std::atof(argv[11]) : 0.1) : 0.0;
  cab_net *net = nullptr;
  if (cab_net_load_text(argv[1], device, &net) != CAB_OK) { std::fprintf(stderr, "%s\n", cab_last_error()); return 1; }

  std::vector<Worker> workers(threads);
  for (int t = 0; t < threads; ++t) {
    Worker &w = workers[t];
    w.net = net;
    w.playouts_per_move = playouts;
    w.moves_per_game = moves;
    w.in_flight = in_flight;
    w.network_priors = network_priors;
    w.save_ratio = save_ratio;
    w.precision = precision;
    w.roots = roots;
    w.rng.seed(seed + 7919ull * t);
    const int lo = games * t / threads, hi = games * (t + 1) / threads;
    w.games.resize(hi - lo);
    for (auto &g : w.games) start_position(g.pos);
  }
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (auto &w : workers) pool.emplace_back([&w] { w.run(); });
  for (auto &th : pool) th.join();
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  Stats tot;
  for (auto &w : workers) {
    tot.playouts += w.st.playouts; tot.boards += w.st.boards; tot.flushes += w.st.flushes;
    tot.expansions += w.st.expansions; tot.moves += w.st.moves; tot.max_flush = std::max(tot.max_flush, w.st.max_flush);
    tot.flush_seconds += w.st.flush_seconds;
  }
  long n_cases = 0;
  if (cases_path)
    for (auto &w : workers) {
      if (!cases::append_shard(cases_path, w.finished_cases)) { std::fprintf(stderr, "cannot append to %s\n", cases_path); return 1; }
      n_cases += (long) w.finished_cases.size();
    }
  long launches = 0;
  cab_net_launch_count(net, &launches);
  std::printf("{\"config\": \"MCTS self-play, virtual-loss leaf batching\", \"games\": %d, \"playouts_per_move\": %d, \"moves_per_game\": %d, "
              "\"threads\": %d, \"in_flight_per_thread\": %d, \"network_priors\": %d, \"seconds\": %.3f, \"moves_played\": %ld, "
              "\"playouts\": %ld, \"playouts_per_s\": %.1f, \"boards_evaluated\": %ld, \"boards_per_s\": %.1f, \"expansions\": %ld, "
              "\"flushes\": %ld, \"avg_boards_per_flush\": %.1f, \"max_boards_per_flush\": %ld, \"gpu_launches\": %ld, \"cases_written\": %ld, \"gpu_wait_fraction\": %.3f, \"precision\": \"%s\", \"roots\": %d}\n",
              games, playouts, moves, threads, in_flight, (int) network_priors, secs, tot.moves, tot.playouts, tot.playouts / secs,
              tot.boards, tot.boards / secs, tot.expansions, tot.flushes, tot.flushes ? (double) tot.boards / tot.flushes : 0.0,
              tot.max_flush, launches, n_cases, tot.flush_seconds / (secs * threads), precision ? "f16" : "fp64", roots);
  cab_net_destroy(net);
  return 0;
}
