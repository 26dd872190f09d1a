// fastchess_c.cpp — plain-C entry points over fastchess.hpp for the Python tests (ctypes). Host-side only.
#include "cases.hpp"
#include "fastchess.hpp"
#include "../csrc/chess_rules.cuh"   // the device rules, compiled for the host: pinned on the CPU by tests/test_chess_rules.py

using namespace fastchess;

extern "C" {
__attribute__((visibility("default"))) void fc_start(Position *p) { start_position(*p); }
__attribute__((visibility("default"))) void fc_set(Position *p, const int8_t *codes, int side, int no_capture) {
  std::memset(p, 0, sizeof *p);
  std::memcpy(p->sq, codes, 64);
  p->side = (int8_t) side;
  p->no_capture = no_capture;
}
// moves_out: 3 bytes per move (from, to, promo); returns the count
__attribute__((visibility("default"))) int fc_moves(const Position *p, uint8_t *moves_out, int queen_only) {
  Move mv[kMaxMoves];
  const int n = generate_moves(*p, p->side, mv, queen_only != 0);
  for (int i = 0; i < n; ++i) { moves_out[3 * i] = mv[i].from; moves_out[3 * i + 1] = mv[i].to; moves_out[3 * i + 2] = mv[i].promo; }
  return n;
}
__attribute__((visibility("default"))) void fc_play(Position *p, int from, int to, int promo) {
  play(*p, Move{(uint8_t) from, (uint8_t) to, (uint8_t) promo});
}
__attribute__((visibility("default"))) long fc_perft(const Position *p, int depth) { return perft(*p, depth); }
__attribute__((visibility("default"))) double fc_material(const Position *p, int color) { return material(*p, color); }
__attribute__((visibility("default"))) int fc_sizeof_position(void) { return (int) sizeof(Position); }

// ---- training cases (cases.hpp) ----
#define FC_API __attribute__((visibility("default")))
FC_API void fc_case_target(double node_value, int game_result, double *bounded_out, double *target_out) {
  *bounded_out = cases::bounded_mcts_result(node_value);
  *target_out = cases::overall_result(*bounded_out, (double) game_result);
}
FC_API int fc_case_text(const int8_t *codes, double target, char *out, int cap) {
  cases::Case c;
  std::memcpy(c.codes, codes, 64);
  c.target = target;
  const std::string s = cases::case_text(c);
  if ((int) s.size() + 1 > cap) return -1;
  std::memcpy(out, s.c_str(), s.size() + 1);
  return (int) s.size();
}
FC_API int fc_case_parse(const char *text, int8_t *codes_out, double *target_out) {
  cases::Case c;
  if (!cases::parse_case_text(text, c)) return 1;
  std::memcpy(codes_out, c.codes, 64);
  *target_out = c.target;
  return 0;
}
FC_API int fc_shard_append(const char *path, const int8_t *codes, const double *targets, long n) {
  std::vector<cases::Case> cs((size_t) n);
  for (long i = 0; i < n; ++i) { std::memcpy(cs[(size_t) i].codes, codes + i * 64, 64); cs[(size_t) i].target = targets[i]; }
  return cases::append_shard(path, cs) ? 0 : 1;
}
// returns the number of cases in the shard (-1 on error); fills at most cap of them
FC_API long fc_shard_read(const char *path, int8_t *codes_out, double *targets_out, long cap) {
  std::vector<cases::Case> cs;
  if (!cases::read_shard(path, cs)) return -1;
  for (long i = 0; i < (long) cs.size() && i < cap; ++i) { std::memcpy(codes_out + i * 64, cs[(size_t) i].codes, 64); targets_out[i] = cs[(size_t) i].target; }
  return (long) cs.size();
}
}

// ---- csrc/chess_rules.cuh (what the MCTS kernels run) through the same kind of entry points ----
extern "C" {
FC_API int cr_moves(const int8_t *codes, int side, uint8_t *moves_out, int queen_only) {
  return crules::generate_moves(codes, side, moves_out, queen_only != 0);
}
FC_API int cr_apply(int8_t *codes, int from, int to, int promo) { return crules::apply_move(codes, from, to, promo) ? 1 : 0; }
FC_API int cr_has_move(const int8_t *codes, int side) { return crules::has_legal_move(codes, side) ? 1 : 0; }
FC_API int cr_material(const int8_t *codes, int color) { return crules::material(codes, color); }
FC_API int cr_in_check(const int8_t *codes, int side) {
  const int ks = crules::king_square(codes, side);
  return ks >= 0 && crules::attacked(crules::plain(codes), ks >> 3, ks & 7, side) ? 1 : 0;
}
}
