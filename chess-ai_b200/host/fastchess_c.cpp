// fastchess_c.cpp — plain-C entry points over fastchess.hpp for the Python tests (ctypes). Host-side only.
#include "fastchess.hpp"

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
}
