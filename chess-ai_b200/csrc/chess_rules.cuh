// chess_rules.cuh — the reference's chess rules on the network's own code board, as __host__ __device__ functions.
//
// Device-side twin of host/fastchess.hpp (same rules, same quirks, same move ORDER; every item cites the reference lines
// it restates) for the MCTS kernels of cab_mcts.cu: node expansion, "has the opponent a reply" (decider.cpp:47-52), the
// moves played during a playout and the end-of-game test all run on the GPU. Plain scalar code over a 64-byte board
// held in shared or local memory; the warp-parallel pieces (one lane per from-square, prefix sums) live in cab_mcts.cu.
// The same header compiles for the host (tests/test_chess_rules.py drives it through libfastchess.so) so the device
// rules are pinned against the recorded reference games on the CPU, and again on the GPU through cab_movegen_host.
//
// Codes k: 1 K, 2 K moved, 3 Q, 4 R, 5 R moved, 6 N, 7 B, 8 P, 9 P that just moved two; black negative; 0 empty;
// square = row * 8 + col, row 0 = white back rank.
//   piece move patterns   /root/reference/src/chess/piece.cpp:110-252
//   castling              piece.cpp:116-136 (only the middle square must be empty; rook keeps its flag, game.cpp:565-576)
//   en passant            piece.cpp:232-244 (any pawn beside with the moved2x flag; flags cleared every move, game.cpp:505-513)
//   legality              Move::verify game.cpp:455-474 + Board::canPieceMove :196-226 (moves ONLY the piece itself)
//   attacks               Board::getPositionThreats game.cpp:66-174
//   move order            game.cpp:228-263 (from ascending, to ascending, promotions Q R N B :404-417)
//   doing a move          Move::doMove game.cpp:499-623
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CR_HD __host__ __device__ __forceinline__
#else
#define CR_HD inline
#endif

namespace crules {

CR_HD int sgn(int x) { return (x > 0) - (x < 0); }
CR_HD int iabs(int x) { return x < 0 ? -x : x; }
CR_HD int color_of(int k) { return sgn(k); }
// 0 none, 1 king, 2 queen, 3 rook, 4 knight, 5 bishop, 6 pawn (nibble a of the constant = kind of |k| = a)
CR_HD int kind_of(int k) {
  const int a = iabs(k);
  return a <= 9 ? (int) ((0x6654332110ull >> (4 * a)) & 15ull) : 0;
}
CR_HD bool on_board(int r, int c) { return (unsigned) r < 8u && (unsigned) c < 8u; }

// The board as Board::canPieceMove sees it while it tries a move: the piece lifted from `from` and put on `to`,
// NOTHING else changed (no en-passant removal, no castling rook; game.cpp:196-226). from = to = -1: the board as it is.
struct View {
  const int8_t *sq;
  int from, to, piece;
  CR_HD int at(int s) const { return s == from ? 0 : (s == to ? piece : (int) sq[s]); }
};
CR_HD View plain(const int8_t *sq) { return View{sq, -1, -1, 0}; }

// Board::getPositionThreats (game.cpp:66-174) > 0 for the square (r, c) of a piece of colour king_color
CR_HD bool attacked(const View &v, int r, int c, int king_color) {
  const int enemy = -king_color;
#pragma unroll 1
  for (int d = 0; d < 8; ++d) {   // 0-3 rank / file, 4-7 diagonals
    const int dr = d < 4 ? (d == 0) - (d == 1) : ((d & 2) ? -1 : 1);
    const int dc = d < 4 ? (d == 2) - (d == 3) : ((d & 1) ? -1 : 1);
    const int slider = d < 4 ? 3 : 5;   // rook / bishop; the queen counts on both
    int x = r + dr, y = c + dc;
    while (on_board(x, y)) {
      const int k = v.at(x * 8 + y);
      if (k != 0) {
        if (color_of(k) == enemy) { const int t = kind_of(k); if (t == 2 || t == slider) return true; }
        break;  // any piece blocks
      }
      x += dr; y += dc;
    }
  }
  // (the short loops below stay rolled on the device: this function is inlined several times per kernel and the search
  // kernels are bound by instruction fetch, not by these few iterations)
#pragma unroll 1
  for (int x = r - 1; x <= r + 1; ++x)
#pragma unroll 1
    for (int y = c - 1; y <= c + 1; ++y)
      if (on_board(x, y)) { const int k = v.at(x * 8 + y); if (color_of(k) == enemy && kind_of(k) == 1) return true; }
  const int px = r + (enemy > 0 ? -1 : 1);  // a white pawn attacks from below, a black one from above
#pragma unroll 1
  for (int dc = -1; dc <= 1; dc += 2)
    if (on_board(px, c + dc)) { const int k = v.at(px * 8 + c + dc); if (color_of(k) == enemy && kind_of(k) == 6) return true; }
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {   // knight jumps (+-1, +-2), (+-2, +-1)
    const int a = (i & 4) ? 2 : 1, b = 3 - a;
    const int x = r + ((i & 1) ? -a : a), y = c + ((i & 2) ? -b : b);
    if (on_board(x, y)) { const int k = v.at(x * 8 + y); if (color_of(k) == enemy && kind_of(k) == 4) return true; }
  }
  return false;
}

// Board::getKingPosition (game.cpp:176-194): first king of that colour in scan order, -1 if none
CR_HD int king_square(const int8_t *sq, int color) {
  for (int i = 0; i < 64; ++i)
    if (kind_of(sq[i]) == 1 && color_of(sq[i]) == color) return i;
  return -1;
}

CR_HD bool aligned(int a, int b) {
  const int dr = (a >> 3) - (b >> 3), dc = (a & 7) - (b & 7);
  return dr == 0 || dc == 0 || dr == dc || dr == -dc;
}

// pseudo-legal targets of the piece on `from` as a 64-bit mask (the piece's verifyMove + "not onto an own piece")
CR_HD uint64_t targets(const int8_t *sq, int from) {
  const int k = sq[from], color = color_of(k), r1 = from >> 3, c1 = from & 7;
  const int kind = kind_of(k);
  uint64_t m = 0;
  if (kind == 2 || kind == 3 || kind == 5) {   // sliders: queen all 8 rays, rook 0-3, bishop 4-7
    const int d0 = kind == 5 ? 4 : 0, d1 = kind == 3 ? 4 : 8;
    for (int d = d0; d < d1; ++d) {
      const int dr = d < 4 ? (d == 0) - (d == 1) : ((d & 2) ? -1 : 1);
      const int dc = d < 4 ? (d == 2) - (d == 3) : ((d & 1) ? -1 : 1);
      int r = r1 + dr, c = c1 + dc;
      while (on_board(r, c)) {
        const int t = sq[r * 8 + c];
        if (color_of(t) != color) m |= 1ull << (r * 8 + c);
        if (t != 0) break;
        r += dr; c += dc;
      }
    }
  } else if (kind == 1) {   // king (piece.cpp:110-137)
    for (int dr = -1; dr <= 1; ++dr)
      for (int dc = -1; dc <= 1; ++dc)
        if ((dr || dc) && on_board(r1 + dr, c1 + dc) && color_of(sq[(r1 + dr) * 8 + c1 + dc]) != color)
          m |= 1ull << ((r1 + dr) * 8 + c1 + dc);
    if (iabs(k) == 1) {     // unmoved: castling
      const View v = plain(sq);
      // rolled loops: ONE inlined copy of attacked() here instead of six (the walk / expand kernels starve on
      // instruction fetch: ncu "no_instruction" 2.6 / 12.5 warps per issue)
#pragma unroll 1
      for (int dir = -1; dir <= 1; dir += 2) {
        const int c2 = c1 + 2 * dir, rook_col = dir > 0 ? 7 : 0, mid = c1 + dir;
        if (!on_board(r1, c2)) continue;
        const int rk = sq[r1 * 8 + rook_col];
        if (!(kind_of(rk) == 3 && color_of(rk) == color && iabs(rk) == 4)) continue;   // own, unmoved rook
        if (color_of(sq[r1 * 8 + c2]) == color) continue;                               // Move::verify
        if (sq[r1 * 8 + mid] != 0) continue;
        bool hit = false;                                                               // king, middle and target square unattacked
#pragma unroll 1
        for (int j = 0; j < 3 && !hit; ++j) hit = attacked(v, r1, c1 + j * dir, color);
        if (hit) continue;
        m |= 1ull << (r1 * 8 + c2);
      }
    }
  } else if (kind == 4) {   // knight
    for (int i = 0; i < 8; ++i) {
      const int a = (i & 4) ? 2 : 1, b = 3 - a;
      const int r = r1 + ((i & 1) ? -a : a), c = c1 + ((i & 2) ? -b : b);
      if (on_board(r, c) && color_of(sq[r * 8 + c]) != color) m |= 1ull << (r * 8 + c);
    }
  } else if (kind == 6) {   // pawn (piece.cpp:210-247)
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
        if (kind_of(beside) == 6 && iabs(beside) == 9) m |= 1ull << (r2 * 8 + c2);   // any pawn flagged moved2x
      }
    }
  }
  return m;
}

CR_HD int ctz64(uint64_t m) {
#if defined(__CUDA_ARCH__)
  return __ffsll((long long) m) - 1;
#else
  return __builtin_ctzll(m);
#endif
}
CR_HD int popc64(uint64_t m) {
#if defined(__CUDA_ARCH__)
  return __popcll(m);
#else
  return __builtin_popcountll(m);
#endif
}

// legal targets of the piece on `from` (colour `color`): the pseudo-legal ones that leave the own king unattacked in
// canPieceMove's simulation. ks = the own king's square (-1: none), calm = that king is not in check now.
// A non-king piece that does not share a rank, file or diagonal with a calm king cannot uncover an attack: it passes
// without the simulation (fastchess.hpp:needs_simulation). A king move is simulated with the king on `to`
// (one king per colour: the reference's representation invariant).
CR_HD uint64_t legal_targets(const int8_t *sq, int from, int color, int ks, bool calm) {
  uint64_t m = targets(sq, from);
  const bool king = kind_of(sq[from]) == 1;
  if (m == 0 || ks < 0 || (calm && !king && !aligned(from, ks))) return m;
  uint64_t legal = 0;
  const int piece = sq[from];
  while (m) {
    const int to = ctz64(m);
    m &= m - 1;
    const View v{sq, from, to, piece};
    const int k2 = king ? to : ks;
    if (!attacked(v, k2 >> 3, k2 & 7, color)) legal |= 1ull << to;
  }
  return legal;
}

// does `color` have any legal move? (Game::updateGameState game.cpp:872-890, decider.cpp:47-52)
CR_HD bool has_legal_move(const int8_t *sq, int color) {
  const int ks = king_square(sq, color);
  const bool calm = ks >= 0 && !attacked(plain(sq), ks >> 3, ks & 7, color);
  for (int from = 0; from < 64; ++from) {
    const int k = sq[from];
    if (k == 0 || color_of(k) != color) continue;
    if (legal_targets(sq, from, color, ks, calm) != 0) return true;
  }
  return false;
}

// Move::doMove (game.cpp:499-623) on the codes; promo = code base of the new piece (3 Q, 4 R, 6 N, 7 B) or 0.
// Returns true when a piece was captured.
// apply_move after the moved2x flags have been cleared (the MCTS walk clears them with all lanes of the warp)
CR_HD bool apply_move_cleared(int8_t *sq, int from, int to, int promo);
CR_HD bool apply_move(int8_t *sq, int from, int to, int promo) {
  for (int i = 0; i < 64; ++i) {                      // every moved2x flag falls (game.cpp:505-513)
    if (sq[i] == 9) sq[i] = 8;
    else if (sq[i] == -9) sq[i] = -8;
  }
  return apply_move_cleared(sq, from, to, promo);
}
CR_HD bool apply_move_cleared(int8_t *sq, int from, int to, int promo) {
  const int r1 = from >> 3, c1 = from & 7, r2 = to >> 3, c2 = to & 7;
  const int k = sq[from], color = color_of(k), kind = kind_of(k);
  bool capture = sq[to] != 0;
  sq[to] = (int8_t) k;
  sq[from] = 0;
  if (kind == 3) {
    if (iabs(k) == 4) sq[to] = (int8_t) (5 * color);                                   // rook has moved (:560-564)
  } else if (kind == 1) {
    if (iabs(k) == 1) sq[to] = (int8_t) (2 * color);                                   // king has moved (:565-570)
    if (iabs(c1 - c2) == 2) {                                                          // castling (:583-595)
      const int rook_col = c2 > c1 ? 7 : 0, mid = (c1 + c2) / 2;
      sq[r2 * 8 + mid] = sq[r1 * 8 + rook_col];                                        // the rook keeps its flag
      sq[r1 * 8 + rook_col] = 0;
    }
  } else if (kind == 6) {
    if (iabs(r1 - r2) == 2) sq[to] = (int8_t) (9 * color);                             // moved2x (:571-576)
    if (r2 == 0 || r2 == 7) sq[to] = (int8_t) (promo * color);                         // new, unmoved piece (:597-622)
    if (iabs(c1 - c2) == 1 && !capture) {                                              // en passant (:624-629)
      capture = sq[r1 * 8 + c2] != 0;
      sq[r1 * 8 + c2] = 0;
    }
  }
  return capture;
}

// material sum the reference's search uses as leaf value (tree.cpp:244-250, piece.fwd.h:236-260), from `color`'s view
CR_HD int material(const int8_t *sq, int color) {
  int s = 0;
  for (int i = 0; i < 64; ++i) {
    const int kd = kind_of(sq[i]);
    const int v = kd == 1 ? 100 : (int) ((0x1335900ull >> (4 * kd)) & 15ull);   // - K Q R N B P = 0 100 9 5 3 3 1
    s += color_of(sq[i]) * v;
  }
  return color * s;
}

// all legal moves of `color` in the reference's generation order; 3 bytes per move (from, to, promo). Scalar; the MCTS
// kernel spreads the from-squares over a warp instead, this form serves tests and single positions.
CR_HD int generate_moves(const int8_t *sq, int color, uint8_t *out, bool queen_only) {
  int n = 0;
  const int ks = king_square(sq, color);
  const bool calm = ks >= 0 && !attacked(plain(sq), ks >> 3, ks & 7, color);
  for (int from = 0; from < 64; ++from) {
    const int k = sq[from];
    if (k == 0 || color_of(k) != color) continue;
    uint64_t m = legal_targets(sq, from, color, ks, calm);
    const bool pawn = kind_of(k) == 6;
    while (m) {
      const int to = ctz64(m);
      m &= m - 1;
      const bool promotes = pawn && ((to >> 3) == 0 || (to >> 3) == 7);
      const int reps = promotes && !queen_only ? 4 : 1;
      for (int p = 0; p < reps; ++p) {
        out[3 * n] = (uint8_t) from;
        out[3 * n + 1] = (uint8_t) to;
        out[3 * n + 2] = (uint8_t) (promotes ? (int) ((0x7643u >> (4 * p)) & 15u) : 0);   // Q R N B = 3 4 6 7
        ++n;
      }
    }
  }
  return n;
}

}  // namespace crules
