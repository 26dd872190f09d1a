// cab_prediction.cu — Decider::prediction (/root/reference/src/mcts_network/decider.cpp:31-61) as one batched call.
//
// The reference evaluates the position, then does / undoes every legal move and evaluates a child only when the
// opponent has no reply (the shipped, inverted branch; otherwise the logit is the constant cm * 20). Here the legal
// moves come from the rules engine on the network's own code board (host/fastchess.hpp, the reference's rules verified
// move list for move list) and the position plus every child that needs the network go to the GPU in ONE evaluation.
// Map semantics kept: keys are ordered (from, to) ascending (game.h:159-167); the four promotions of a pawn move share
// one key whose value is the LAST one assigned, i.e. the bishop promotion's (actionMap[move] = ..., decider.cpp:52-56).
#include <vector>

#include "../host/fastchess.hpp"
#include "cab_internal.cuh"

using namespace cab;

extern "C" {

int cab_eval_host(cab_net *net, const int8_t *codes_host, long n_boards, double *out_host);

int cab_prediction_host(cab_net *net, const int8_t *codes64, int side_to_move, int network_priors, double *eval_out, int *moves_out,
                        double *logits_out, int cap, int *n_moves) {
  if (!net || !codes64 || !eval_out || !n_moves || (side_to_move != 1 && side_to_move != -1) || cap < 0 ||
      (cap > 0 && (!moves_out || !logits_out))) {
    set_error("bad argument");
    return CAB_EINVAL;
  }
  fastchess::Position p;
  std::memset(&p, 0, sizeof p);
  std::memcpy(p.sq, codes64, 64);
  p.side = (int8_t) side_to_move;
  fastchess::Move mv[fastchess::kMaxMoves];
  const int n = fastchess::generate_moves(p, p.side, mv, /*queen_only=*/false);
  const double cm = side_to_move;

  std::vector<int8_t> boards(64);                     // slot 0: the position itself
  std::memcpy(boards.data(), codes64, 64);
  struct Entry { int from, to, slot; };               // slot < 0: constant logit
  std::vector<Entry> entries;
  for (int i = 0; i < n; ++i) {
    fastchess::Position c = p;
    fastchess::apply_move(c.sq, mv[i]);
    c.side = (int8_t) -p.side;
    bool ask = network_priors != 0;
    if (!ask) {
      fastchess::Move reply[fastchess::kMaxMoves];
      ask = fastchess::generate_moves(c, c.side, reply, true) == 0;    // decider.cpp:47-52
    }
    int slot = -1;
    if (ask) {
      slot = (int) (boards.size() / 64);
      boards.insert(boards.end(), c.sq, c.sq + 64);
    }
    if (!entries.empty() && entries.back().from == mv[i].from && entries.back().to == mv[i].to) entries.back().slot = slot;  // same key: last value wins
    else entries.push_back({mv[i].from, mv[i].to, slot});
  }
  std::vector<double> y(boards.size() / 64);
  const int rc = cab_eval_host(net, boards.data(), (long) y.size(), y.data());
  if (rc != CAB_OK) return rc;
  *eval_out = cm * y[0];
  *n_moves = (int) entries.size();
  for (int i = 0; i < (int) entries.size() && i < cap; ++i) {
    moves_out[2 * i] = entries[i].from;
    moves_out[2 * i + 1] = entries[i].to;
    logits_out[i] = entries[i].slot >= 0 ? cm * y[entries[i].slot] : cm * 20.0;
  }
  return CAB_OK;
}

}  // extern "C"
