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
int cab_eval_f16_host(cab_net *net, const int8_t *codes_host, long n_boards, double *out_host);
int cab_eval_bf16_host(cab_net *net, const int8_t *codes_host, long n_boards, double *out_host);

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


// The move the network ranks first, with the reduced-precision evaluators made SAFE for ranking (BASELINE.json north
// star: "stated bf16 tolerance, with identical argmax move"). With network-driven priors the weight of a move is
// exp(cm * logit) = exp(cm * cm * net(child)) = exp(net(child)) (decider.cpp:52-56, tree.cpp:302), so the first-ranked
// move is the FIRST child in std::map<Move> order with the greatest net(child).
//   precision 0: every child in fp64 — the reference answer.
//   precision 1 (fused fp16 tcgen05, default-shaped nets) / 2 (bf16 tcgen05, wide nets): all children in reduced
//   precision; then every child whose value lies within `tol` of the reduced-precision maximum is RE-EVALUATED in fp64
//   and the maximum is taken over the refined values. If the reduced path errs by at most tol / 2 per board (the
//   stated tolerances: 0.05 fp16, 0.15 bf16 -> default tol 0.1 / 0.3 when tol <= 0), the fp64 winner is provably
//   inside the refined set (its reduced value is >= v* - tol/2, the reduced maximum <= v* + tol/2), so the chosen move
//   is IDENTICAL to the fp64 choice, ties included (same first-wins rule). *n_refined = boards re-evaluated in fp64.
int cab_best_move_host(cab_net *net, const int8_t *codes64, int side_to_move, int precision, double tol, int *from_out, int *to_out,
                       double *value_out, int *n_moves, int *n_refined) {
  if (!net || !codes64 || (side_to_move != 1 && side_to_move != -1) || precision < 0 || precision > 2 || !from_out || !to_out) {
    set_error("bad argument");
    return CAB_EINVAL;
  }
  fastchess::Position p;
  std::memset(&p, 0, sizeof p);
  std::memcpy(p.sq, codes64, 64);
  p.side = (int8_t) side_to_move;
  fastchess::Move mv[fastchess::kMaxMoves];
  const int n = fastchess::generate_moves(p, p.side, mv, /*queen_only=*/false);
  // one board per std::map key: the four promotions of a pawn move share a key whose value is the LAST assigned (bishop)
  std::vector<int8_t> boards;
  std::vector<std::pair<int, int>> keys;
  for (int i = 0; i < n; ++i) {
    fastchess::Position c = p;
    fastchess::apply_move(c.sq, mv[i]);
    if (!keys.empty() && keys.back().first == mv[i].from && keys.back().second == mv[i].to) {
      std::memcpy(boards.data() + (keys.size() - 1) * 64, c.sq, 64);
    } else {
      keys.push_back({mv[i].from, mv[i].to});
      boards.insert(boards.end(), c.sq, c.sq + 64);
    }
  }
  const int k = (int) keys.size();
  if (n_moves) *n_moves = k;
  if (n_refined) *n_refined = 0;
  if (k == 0) { *from_out = *to_out = -1; return CAB_OK; }
  std::vector<double> y(k);
  int rc = precision == 0 ? cab_eval_host(net, boards.data(), k, y.data())
                          : (precision == 1 ? cab_eval_f16_host(net, boards.data(), k, y.data()) : cab_eval_bf16_host(net, boards.data(), k, y.data()));
  if (rc != CAB_OK) return rc;
  if (precision != 0) {
    if (tol <= 0.0) tol = precision == 1 ? 0.1 : 0.3;
    double mx = y[0];
    for (int i = 1; i < k; ++i) mx = y[i] > mx ? y[i] : mx;
    std::vector<int> near;
    std::vector<int8_t> sub;
    for (int i = 0; i < k; ++i)
      if (y[i] >= mx - tol) { near.push_back(i); sub.insert(sub.end(), boards.begin() + (size_t) i * 64, boards.begin() + (size_t) (i + 1) * 64); }
    std::vector<double> exact(near.size());
    if ((rc = cab_eval_host(net, sub.data(), (long) near.size(), exact.data())) != CAB_OK) return rc;
    for (int i = 0; i < k; ++i) y[i] = -1e300;                 // only refined boards compete
    for (size_t j = 0; j < near.size(); ++j) y[near[j]] = exact[j];
    if (n_refined) *n_refined = (int) near.size();
  }
  int best = 0;
  for (int i = 1; i < k; ++i)
    if (y[i] > y[best]) best = i;                              // strictly greater: the first maximum wins (tree.cpp:108)
  *from_out = keys[best].first;
  *to_out = keys[best].second;
  if (value_out) *value_out = y[best];
  return CAB_OK;
}

}  // extern "C"
