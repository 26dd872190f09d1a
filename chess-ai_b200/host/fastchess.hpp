// fastchess.hpp — fast host-side restatement of the reference's chess rules ON THE NETWORK'S OWN BOARD FORMAT
// (SURVEY.md §8f N2). Header-only C++17, no dependencies.
//
// A position is the 64 int8 codes the evaluator consumes (k: 1 K, 2 K moved, 3 Q, 4 R, 5 R moved, 6 N, 7 B, 8 P,
// 9 P that just moved two; black negative; 0 empty; square = row*8+col, row 0 = white back rank), plus the side to
// move and the no-capture counter — so a searched position goes to the GPU with no encoding step at all.
//
// The rules are the REFERENCE's, quirks included (paths under /root/reference/src/chess/), because the search must
// walk the same tree:
//   piece move patterns            piece.cpp:110-252 (verifyMove per piece, checkClearMovePath :80-91)
//   castling                       piece.cpp:116-136: king and that-side rook unmoved, king's start, end and middle
//                                  squares unattacked, ONLY the middle square must be empty; the rook keeps its
//                                  "unmoved" flag (game.cpp:565-576 flags the king only); an enemy piece on the king's
//                                  target square is captured like by any king move
//   en passant                     piece.cpp:232-244: capturing pawn on its 5th rank, ANY pawn beside it whose
//                                  moved2x flag is set; every moved2x flag is cleared at the start of each move
//                                  (game.cpp:505-513), so only the previous mover's pawn qualifies
//   legality                       Move::verify game.cpp:455-474 + Board::canPieceMove :196-226: the simulation moves
//                                  ONLY the piece itself (no en-passant removal, no castling rook) and asks whether
//                                  the first own king in scan order is attacked (getKingPosition :176-194)
//   attacks                        Board::getPositionThreats game.cpp:66-174
//   move order                     Board::getPossibleMoves / getMovesFromSquare game.cpp:228-263: from-square
//                                  ascending, to-square ascending, promotions queen, rook, knight, bishop (:404-417)
//   doing a move                   Move::doMove game.cpp:499-623 (a promoted rook is a NEW rook: unmoved)
//   game end                       Game::tryMove :847-871 (50 PLIES without capture = draw, pawn moves do not reset),
//                                  Game::updateGameState :872-890 (no moves: stalemate if the king is safe, else loss);
//                                  no repetition rule
//   leaf value of the search       tree.cpp:244-253 material sum with PieceType::minimaxValue piece.fwd.h:236-260
// Verified move list by move list and board by board against games recorded from the reference's own code
// (tests/golden/chess_games.npz, tests/test_fastchess.py) and by perft 20 / 400 / 8 902.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace fastchess {

struct Move {
  uint8_t from, to, promo;  // promo: 0 none, else the code base of the new piece (3 Q, 4 R, 6 N, 7 B)
  bool operator==(const Move &o) const { return from == o.from && to == o.to && promo == o.promo; }
};

struct Position {
  int8_t sq[64];
  int8_t side;        // +1 white to move, -1 black
  int8_t over;        // game finished
  int8_t result;      // GameResult::evaluate(): +1 white won, -1 black won, 0 draw / undecided
  int8_t pad;
  int32_t no_capture; // plies since the last capture
};

constexpr int kMaxMoves = 256;

inline int sgn(int x) { return (x > 0) - (x < 0); }
inline int color_of(int k) { return sgn(k); }
inline int kind_of(int k) {  // 0 none, 1 king, 2 queen, 3 rook, 4 knight, 5 bishop, 6 pawn
  static const int8_t t[10] = {0, 1, 1, 2, 3, 3, 4, 5, 6, 6};
  const int a = std::abs(k);
  return a <= 9 ? t[a] : 0;
}
inline bool on_board(int r, int c) { return (unsigned) r < 8u && (unsigned) c < 8u; }

inline void start_position(Position &p) {
  static const int8_t back[8] = {4, 6, 7, 3, 1, 7, 6, 4};
  std::memset(&p, 0, sizeof p);
  for (int c = 0; c < 8; ++c) {
    p.sq[c] = back[c];
    p.sq[8 + c] = 8;
    p.sq[48 + c] = -8;
    p.sq[56 + c] = (int8_t) -back[c];
  }
  p.side = 1;
}

// Board::getPositionThreats (game.cpp:66-174) > 0
inline bool attacked(const int8_t *sq, int r, int c, int king_color) {
  const int enemy = -king_color;
  static const int axis[4][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}}, diag[4][2] = {{1, 1}, {1, -1}, {-1, 1}, {-1, -1}};
  for (auto &d : axis) {
    int x = r + d[0], y = c + d[1];
    while (on_board(x, y)) {
      const int k = sq[x * 8 + y];
      if (k != 0) {
        if (color_of(k) == enemy) { const int t = kind_of(k); if (t == 2 || t == 3) return true; }
        break;  // any piece blocks
      }
      x += d[0]; y += d[1];
    }
  }
  for (auto &d : diag) {
    int x = r + d[0], y = c + d[1];
    while (on_board(x, y)) {
      const int k = sq[x * 8 + y];
      if (k != 0) {
        if (color_of(k) == enemy) { const int t = kind_of(k); if (t == 2 || t == 5) return true; }
        break;
      }
      x += d[0]; y += d[1];
    }
  }
  for (int x = r - 1; x <= r + 1; ++x)
    for (int y = c - 1; y <= c + 1; ++y)
      if (on_board(x, y)) { const int k = sq[x * 8 + y]; if (color_of(k) == enemy && kind_of(k) == 1) return true; }
  const int px = r + (enemy > 0 ? -1 : 1);  // a white pawn attacks from below, a black one from above
  for (int dc = -1; dc <= 1; dc += 2)
    if (on_board(px, c + dc)) { const int k = sq[px * 8 + c + dc]; if (color_of(k) == enemy && kind_of(k) == 6) return true; }
  static const int kn[8][2] = {{1, 2}, {1, -2}, {-1, 2}, {-1, -2}, {2, 1}, {2, -1}, {-2, 1}, {-2, -1}};
  for (auto &d : kn)
    if (on_board(r + d[0], c + d[1])) { const int k = sq[(r + d[0]) * 8 + c + d[1]]; if (color_of(k) == enemy && kind_of(k) == 4) return true; }
  return false;
}

// Board::getKingPosition (game.cpp:176-194): first king of that colour in scan order
inline int king_square(const int8_t *sq, int color) {
  for (int i = 0; i < 64; ++i)
    if (kind_of(sq[i]) == 1 && color_of(sq[i]) == color) return i;
  return -1;
}

// Board::canPieceMove (game.cpp:196-226): move ONLY the piece, then ask whether the own king is attacked
inline bool leaves_king_safe(const Position &p, int from, int to) {
  int8_t tmp[64];
  std::memcpy(tmp, p.sq, 64);
  const int color = color_of(tmp[from]);
  tmp[to] = tmp[from];
  tmp[from] = 0;
  const int ks = king_square(tmp, color);
  if (ks < 0) return true;  // no king of that colour on the board (never happens from a legal start)
  return !attacked(tmp, ks >> 3, ks & 7, color);
}

// The same predicate for a NON-king piece when the king (on ks, not in check now) is known: moving a piece can only
// uncover a sliding attack, and only through the square it leaves; landing on a square or capturing can block or
// remove attackers but never add one. So a piece that does not share a rank, file or diagonal with the king passes
// without the simulation; everything else (aligned pieces, any move while in check, king moves) is simulated.
inline bool leaves_king_safe_known_king(const Position &p, int from, int to, int ks, int color) {   // non-king piece
  int8_t tmp[64];
  std::memcpy(tmp, p.sq, 64);
  tmp[to] = tmp[from];
  tmp[from] = 0;
  return !attacked(tmp, ks >> 3, ks & 7, color);
}
inline bool aligned(int a, int b) {
  const int dr = (a >> 3) - (b >> 3), dc = (a & 7) - (b & 7);
  return dr == 0 || dc == 0 || dr == dc || dr == -dc;
}

// pseudo-legal targets of the piece on `from` as a 64-bit mask (the piece's verifyMove + "not onto an own piece")
inline uint64_t targets(const Position &p, int from) {
  const int8_t *sq = p.sq;
  const int k = sq[from], color = color_of(k), r1 = from >> 3, c1 = from & 7;
  uint64_t m = 0;
  auto add = [&](int r, int c) { if (on_board(r, c) && color_of(sq[r * 8 + c]) != color) m |= 1ull << (r * 8 + c); };
  auto ray = [&](int dr, int dc) {
    int r = r1 + dr, c = c1 + dc;
    while (on_board(r, c)) {
      const int t = sq[r * 8 + c];
      if (color_of(t) != color) m |= 1ull << (r * 8 + c);
      if (t != 0) break;
      r += dr; c += dc;
    }
  };
  switch (kind_of(k)) {
    case 1: {  // king (piece.cpp:110-137)
      for (int dr = -1; dr <= 1; ++dr)
        for (int dc = -1; dc <= 1; ++dc)
          if (dr || dc) add(r1 + dr, c1 + dc);
      if (std::abs(k) == 1) {  // unmoved
        for (int dir = -1; dir <= 1; dir += 2) {
          const int c2 = c1 + 2 * dir, rook_col = dir > 0 ? 7 : 0, mid = c1 + dir;
          if (!on_board(r1, c2)) continue;
          const int rk = sq[r1 * 8 + rook_col];
          if (!(kind_of(rk) == 3 && color_of(rk) == color && std::abs(rk) == 4)) continue;  // own, unmoved rook
          if (color_of(sq[r1 * 8 + c2]) == color) continue;                                  // Move::verify
          if (attacked(sq, r1, c1, color) || attacked(sq, r1, c2, color)) continue;
          if (sq[r1 * 8 + mid] != 0 || attacked(sq, r1, mid, color)) continue;
          m |= 1ull << (r1 * 8 + c2);
        }
      }
      break;
    }
    case 2: ray(1, 0); ray(-1, 0); ray(0, 1); ray(0, -1); ray(1, 1); ray(1, -1); ray(-1, 1); ray(-1, -1); break;
    case 3: ray(1, 0); ray(-1, 0); ray(0, 1); ray(0, -1); break;
    case 5: ray(1, 1); ray(1, -1); ray(-1, 1); ray(-1, -1); break;
    case 4: {
      static const int kn[8][2] = {{1, 2}, {1, -2}, {-1, 2}, {-1, -2}, {2, 1}, {2, -1}, {-2, 1}, {-2, -1}};
      for (auto &d : kn) add(r1 + d[0], c1 + d[1]);
      break;
    }
    case 6: {  // pawn (piece.cpp:210-247)
      const int dir = color > 0 ? 1 : -1, home = color > 0 ? 1 : 6, r2 = r1 + dir;
      if (on_board(r2, c1) && sq[r2 * 8 + c1] == 0) {
        m |= 1ull << (r2 * 8 + c1);
        if (r1 == home && sq[(r1 + 2 * dir) * 8 + c1] == 0) m |= 1ull << ((r1 + 2 * dir) * 8 + c1);
      }
      for (int dc = -1; dc <= 1; dc += 2) {
        const int c2 = c1 + dc;
        if (!on_board(r2, c2)) continue;
        const int t = sq[r2 * 8 + c2];
        if (t != 0) { if (color_of(t) != color) m |= 1ull << (r2 * 8 + c2); }
        else if (r1 == home + 3 * dir) {
          const int beside = sq[r1 * 8 + c2];
          if (kind_of(beside) == 6 && std::abs(beside) == 9) m |= 1ull << (r2 * 8 + c2);  // any pawn flagged moved2x
        }
      }
      break;
    }
    default: break;
  }
  return m;
}

// legal moves of `color` in the reference's generation order; all four promotions unless queen_only
inline int generate_moves(const Position &p, int color, Move *out, bool queen_only = false) {
  int n = 0;
  const int ks = king_square(p.sq, color);
  const bool calm = ks >= 0 && !attacked(p.sq, ks >> 3, ks & 7, color);   // own king present and not in check
  for (int from = 0; from < 64; ++from) {
    const int k = p.sq[from];
    if (k == 0 || color_of(k) != color) continue;
    uint64_t m = targets(p, from);
    const bool pawn = kind_of(k) == 6;
    const bool king = kind_of(k) == 1;
    const bool needs_simulation = !(calm && !king && !aligned(from, ks));
    while (m) {
      const int to = __builtin_ctzll(m);
      m &= m - 1;
      // a non-king move never moves or captures an own king, so the scan inside leaves_king_safe would find ks again
      if (needs_simulation && !(king || ks < 0 ? leaves_king_safe(p, from, to) : leaves_king_safe_known_king(p, from, to, ks, color))) continue;
      if (pawn && ((to >> 3) == 0 || (to >> 3) == 7)) {
        out[n++] = Move{(uint8_t) from, (uint8_t) to, 3};
        if (!queen_only) {
          out[n++] = Move{(uint8_t) from, (uint8_t) to, 4};
          out[n++] = Move{(uint8_t) from, (uint8_t) to, 6};
          out[n++] = Move{(uint8_t) from, (uint8_t) to, 7};
        }
      } else out[n++] = Move{(uint8_t) from, (uint8_t) to, 0};
    }
  }
  return n;
}

// Move::doMove (game.cpp:499-623) on the codes; returns true when a piece was captured
// every moved2x flag falls (game.cpp:505-513): codes 9 -> 8 and -9 -> -8 on all 64 squares, eight squares per word.
// z has 0x80 in every byte of v that equals the pattern byte (exact zero-byte test, no cross-byte carries).
inline void clear_moved2x(int8_t *sq) {
  constexpr uint64_t k7f = 0x7f7f7f7f7f7f7f7full, k01 = 0x0101010101010101ull;
  for (int i = 0; i < 8; ++i) {
    uint64_t v;
    std::memcpy(&v, sq + 8 * i, 8);
    const uint64_t a = v ^ (k01 * 0x09), b = v ^ (k01 * 0xf7);
    const uint64_t za = ~(((a & k7f) + k7f) | a | k7f), zb = ~(((b & k7f) + k7f) | b | k7f);
    if ((za | zb) == 0) continue;
    v = v - (za >> 7) + (zb >> 7);     // 0x09 -> 0x08, 0xf7 -> 0xf8: no borrow or carry leaves the byte
    std::memcpy(sq + 8 * i, &v, 8);
  }
}

inline bool apply_move(int8_t *sq, Move mv) {
  clear_moved2x(sq);
  const int from = mv.from, to = mv.to, r1 = from >> 3, c1 = from & 7, r2 = to >> 3, c2 = to & 7;
  int k = sq[from];
  const int color = color_of(k);
  bool capture = sq[to] != 0;
  sq[to] = (int8_t) k;
  sq[from] = 0;
  switch (kind_of(k)) {
    case 3: if (std::abs(k) == 4) sq[to] = (int8_t) (5 * color); break;                 // rook has moved (:560-564)
    case 1:
      if (std::abs(k) == 1) sq[to] = (int8_t) (2 * color);                              // king has moved (:565-570)
      if (std::abs(c1 - c2) == 2) {                                                     // castling (:583-595)
        const int rook_col = c2 > c1 ? 7 : 0, mid = (c1 + c2) / 2;
        sq[r2 * 8 + mid] = sq[r1 * 8 + rook_col];                                       // the rook keeps its flag
        sq[r1 * 8 + rook_col] = 0;
      }
      break;
    case 6:
      if (std::abs(r1 - r2) == 2) sq[to] = (int8_t) (9 * color);                        // moved2x (:571-576)
      if (r2 == 0 || r2 == 7) sq[to] = (int8_t) (mv.promo * color);                     // new, unmoved piece (:597-622)
      if (std::abs(c1 - c2) == 1 && !capture) {                                         // en passant (:624-629)
        capture = sq[r1 * 8 + c2] != 0;
        sq[r1 * 8 + c2] = 0;
      }
      break;
    default: break;
  }
  return capture;
}

// Game::updateGameState (game.cpp:872-890)
inline void update_game_state(Position &p) {
  Move tmp[kMaxMoves];
  if (generate_moves(p, p.side, tmp, true) == 0) {
    p.over = 1;
    const int ks = king_square(p.sq, p.side);
    const bool safe = ks < 0 || !attacked(p.sq, ks >> 3, ks & 7, p.side);
    p.result = safe ? 0 : (int8_t) -p.side;
  }
}

// Game::tryMove (game.cpp:847-871) for a move already known to be legal
inline void play(Position &p, Move mv) {
  const bool capture = apply_move(p.sq, mv);
  p.side = (int8_t) -p.side;
  p.no_capture = capture ? 0 : p.no_capture + 1;
  if (p.no_capture >= 50) { p.over = 1; p.result = 0; }
  else update_game_state(p);
}

// material sum the reference's search uses as leaf value (tree.cpp:244-250, piece.fwd.h:236-260), from `color`'s view
inline double material(const Position &p, int color) {
  static const int v[7] = {0, 100, 9, 5, 3, 3, 1};
  int s = 0;
  for (int i = 0; i < 64; ++i) s += color_of(p.sq[i]) * v[kind_of(p.sq[i])];
  return (double) (color * s);
}

inline long perft(const Position &p, int depth) {
  if (depth == 0) return 1;
  if (p.over) return 0;
  Move mv[kMaxMoves];
  const int n = generate_moves(p, p.side, mv);
  long total = 0;
  for (int i = 0; i < n; ++i) {
    Position q = p;
    play(q, mv[i]);
    total += perft(q, depth - 1);
  }
  return total;
}

}  // namespace fastchess
