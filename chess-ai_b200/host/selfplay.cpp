// selfplay.cpp — MCTS self-play with virtual-loss leaf batching (BASELINE.json configs[2]): command-line driver.
//
// The search itself lives on the GPU (csrc/cab_mcts.cu behind cab_mcts_*: trees, playouts, node expansion with the
// reference's rules, priors, move choice, game end, case sampling — tree.cpp / decider.cpp:31-61 / game.cpp:847-890
// restated as kernels and pinned against the reference's own tree::MCTS::mcts(), tests/test_mcts_gpu.py). This file
// only parses the arguments, starts the run, appends the training cases of the finished games to a packed shard
// (host/cases.hpp; make_cases.cpp:36-59) and prints one JSON object. One process per GPU; processes never talk.
// (Round 1's host-side tree walk over per-thread leaf queues is gone: it was bound by the host cores — 0.35 scaling
// efficiency at 8 GPUs — and was not pinned against the reference. The leaf-queue ABI, cab_leafq_*, remains for callers
// that keep the reference's own search, integration/batched_mcts.cpp.)
//
//   selfplay <latest.txt> <games> <playouts_per_move> <moves_per_game> <unused> <in_flight_per_tree> <network_priors>
//            [device] [seed] [cases_path] [save_ratio] [precision: fp64 | f16] [roots] [noise]
// (roots = independent trees per move, merged like the reference's root-parallel threads; each gets playouts / roots;
//  cases_path "-" = none; noise 0 switches the root Dirichlet noise off: deterministic searches)
// No CPU fallback: needs libchessai_b200.so and a GPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/chessai_b200.h"
#include "cases.hpp"

int main(int argc, char **argv) {
  if (argc < 8) {
    std::fprintf(stderr, "usage: %s latest.txt games playouts_per_move moves_per_game - in_flight_per_tree network_priors "
                         "[device] [seed] [cases_path] [save_ratio] [fp64|f16] [roots] [noise]\n", argv[0]);
    return 2;
  }
  cab_mcts_params p;
  std::memset(&p, 0, sizeof p);
  p.games = std::atoi(argv[2]);
  p.playouts_per_move = std::atoi(argv[3]);
  p.max_moves = std::atoi(argv[4]);
  p.in_flight = std::atoi(argv[6]);
  p.network_priors = std::atoi(argv[7]) != 0;
  const int device = argc > 8 ? std::atoi(argv[8]) : 0;
  p.seed = argc > 9 ? std::strtoull(argv[9], nullptr, 10) : 1234;
  const char *cases_path = argc > 10 && std::strcmp(argv[10], "-") != 0 ? argv[10] : nullptr;
  p.save_ratio = cases_path ? (argc > 11 ? std::atof(argv[11]) : 0.1) : 0.0;
  if (argc > 12 && std::strcmp(argv[12], "fp64") != 0 && std::strcmp(argv[12], "f16") != 0) {
    std::fprintf(stderr, "precision '%s': fp64 (the reference's search) or f16 (fused fp16 tcgen05 evaluator, reduced precision)\n", argv[12]);
    return 2;
  }
  p.precision = argc > 12 && std::strcmp(argv[12], "f16") == 0 ? 1 : 0;
  p.roots = argc > 13 ? std::atoi(argv[13]) : 1;
  p.noise = argc > 14 ? std::atoi(argv[14]) != 0 : 1;
  if (p.roots < 1) p.roots = 1;
  if (p.in_flight < 1) p.in_flight = 1;

  cab_net *net = nullptr;
  cab_mcts *q = nullptr;
  if (cab_net_load_text(argv[1], device, &net) != CAB_OK || cab_mcts_create(net, &p, &q) != CAB_OK) {
    std::fprintf(stderr, "%s\n", cab_last_error());
    return 1;
  }
  cab_mcts_stats st;
  if (cab_mcts_run(q, 0, 0, &st) != CAB_OK) { std::fprintf(stderr, "%s\n", cab_last_error()); return 1; }

  long n_cases = 0;
  if (cases_path) {
    if (cab_mcts_cases(q, nullptr, nullptr, 0, &n_cases) != CAB_OK) { std::fprintf(stderr, "%s\n", cab_last_error()); return 1; }
    std::vector<int8_t> codes((size_t) n_cases * 64 + 64);
    std::vector<double> targets((size_t) n_cases + 1);
    if (cab_mcts_cases(q, codes.data(), targets.data(), n_cases, &n_cases) != CAB_OK) { std::fprintf(stderr, "%s\n", cab_last_error()); return 1; }
    std::vector<cases::Case> cs((size_t) n_cases);
    for (long i = 0; i < n_cases; ++i) { std::memcpy(cs[(size_t) i].codes, codes.data() + i * 64, 64); cs[(size_t) i].target = targets[(size_t) i]; }
    if (!cases::append_shard(cases_path, cs)) { std::fprintf(stderr, "cannot append to %s\n", cases_path); return 1; }
  }
  long launches = 0;
  cab_net_launch_count(net, &launches);
  std::printf("{\"config\": \"MCTS self-play resident on the GPU, virtual-loss leaf batching\", \"games\": %d, \"playouts_per_move\": %d, "
              "\"moves_per_game\": %d, \"roots\": %d, \"in_flight_per_tree\": %d, \"network_priors\": %d, \"noise\": %d, \"device\": %d, "
              "\"seconds\": %.4f, \"moves_played\": %ld, \"games_finished\": %ld, \"playouts\": %ld, \"playouts_per_s\": %.1f, "
              "\"boards_evaluated\": %ld, \"boards_per_s\": %.1f, \"expansions\": %ld, \"flushes\": %ld, \"avg_boards_per_flush\": %.1f, "
              "\"evaluator_busy\": %.3f, \"gpu_launches\": %ld, \"cases_written\": %ld, \"arena_full\": %ld, \"eval_full\": %ld, "
              "\"precision\": \"fp64\"}\n",
              p.games, p.playouts_per_move, p.max_moves, p.roots, p.in_flight, p.network_priors, p.noise, device, st.seconds, st.moves_played,
              st.games_finished, st.playouts, st.playouts / st.seconds, st.boards_evaluated, st.boards_evaluated / st.seconds, st.expansions,
              st.rounds, st.rounds ? (double) st.boards_evaluated / st.rounds : 0.0, st.evaluator_seconds / st.seconds, launches, n_cases,
              st.arena_full, st.eval_full);
  cab_mcts_destroy(q);
  cab_net_destroy(net);
  return 0;
}
