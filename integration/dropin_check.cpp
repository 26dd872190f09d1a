// dropin_check.cpp — links the reference's OWN decider / MCTS / player / game code (compiled where it lies under
// /root/reference/src) with integration/network_b200.cpp in place of the reference's network.cpp, and drives the
// reference's public API on the GPU:  NetworkStorage::initialize(path), Decider::prediction(Game*),
// network::load_training_case + Optimizer::optimize, operator<<, tree::MCTS::run_mcts_multithreaded.
// Prints one JSON object; tests/test_dropin_gpu.py compares it with the golden values of the reference's CPU network.
//
//   dropin_check <latest.txt> <start_position.txt> <case.txt> <dump_out.txt> [mcts_sims] [rotate]
// With "rotate" NetworkStorage::SAVE_NETWORKS is on from the start (initialization.cpp:192): initialize() writes
// network_dump/latest.txt, saveNetwork() / flushStorage() rotate it to gen-0.txt, gen-1.txt (network.cpp:345-364); the
// directory network_dump/ must exist under the working directory.
#include <cstdio>
#include <fstream>
#include <string>

#include "chess/game.h"
#include "main/network/make_cases.h"
#include "mcts_network/decider.h"
#include "mcts_network/network.h"
#include "mcts_network/tree.h"

int main(int argc, char **argv) {
  if (argc < 5) { std::fprintf(stderr, "usage: %s latest.txt start.txt case.txt dump_out.txt [mcts_sims]\n", argv[0]); return 2; }
  const bool rotate = argc > 6 && std::string(argv[6]) == "rotate";
  network::NetworkStorage::SAVE_NETWORKS = rotate;
  network::Network *net = network::NetworkStorage::initialize(std::string(argv[1]));

  // Decider::prediction on the start position (decider.cpp:31-61) — reference code calling our predictPosition
  auto *blank = new game::Game(8, 8);
  blank->board()->loadFromFile(argv[2]);
  game::Game *g = blank->clone();  // clone() refreshes the cached move lists
  auto pr = net->prediction(g);
  double min_logit = 1e300, max_logit = -1e300;
  for (auto &kv : pr.second) { min_logit = std::min(min_logit, kv.second); max_logit = std::max(max_logit, kv.second); }

  // one Optimizer::optimize step on a training case (train.cpp:50-51)
  auto tc = network::load_training_case(argv[3], false);
  double lambda = network::Optimizer::DEFAULT_INITIAL_LAMBDA;
  network::Optimizer::optimize(net, tc.first, tc.second, lambda);
  network::Network *copy = net->clone();
  auto pr2 = copy->prediction(g);

  // byte-exact dump through operator<<
  { std::ofstream out(argv[4]); out << copy; }

  if (rotate) network::NetworkStorage::saveNetwork();   // latest.txt -> gen-0.txt, the optimised network -> latest.txt

  // the reference's root-parallel MCTS (tree.cpp:149-196) searching with our evaluator from 2 threads
  int sims = argc > 5 ? std::atoi(argv[5]) : 0;
  int children = -1;
  if (sims > 0) {
    tree::MCTS::NUM_SIMULATIONS_PER_THREAD = sims;
    tree::MCTS::SIMULATION_SEARCH_DEPTH = 3;
    auto mv = tree::MCTS::run_mcts_multithreaded(g, 2, net);
    children = mv.second->countChildren();
    delete mv.second;
  }
  std::printf("{\"dims\": %d, \"eval\": %.17e, \"n_moves\": %zu, \"min_logit\": %.17g, \"max_logit\": %.17g, "
              "\"lambda_after\": %.17e, \"eval_after\": %.17e, \"mcts_root_children\": %d}\n",
              (int) 5, pr.first, pr.second.size(), min_logit, max_logit, lambda, pr2.first, children);
  delete copy;
  delete tc.first;
  network::NetworkStorage::flushStorage();
  return 0;
}
