// encode_simd.cuh — Piece::code()/0.4 for FOUR squares at a time, byte-parallel in 32-bit registers (K0's inner step).
//
// Reference rule (/root/reference/src/chess/piece.cpp:103-105,145-147,179-181,260-262; piece.fwd.h:122-133,287-308):
//   base by type: none 0, king 1, queen 3, rook 4, knight 6, bishop 7, pawn 8; king / rook / pawn add 1 when the
//   moved flag is set; black negates; anything that is not a piece type 1..6 of colour 1..2 encodes as 0.
// The scalar form costs ~37 issued instructions per square (ncu: the kernel was issue-bound at 53 % of the HBM rate);
// here the three 8-entry tables are PRMT byte permutes and the comparisons are carry-free byte tricks: ~10 per square.
// Host-compilable (tests/test_abi_host.py checks it exhaustively against the scalar rule with g++).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define CAB_HD __host__ __device__ __forceinline__
#else
#define CAB_HD inline
#endif

namespace cab {

CAB_HD uint32_t byte_perm(uint32_t x, uint32_t y, uint32_t s) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(x, y, s);
#else
  const uint64_t pool = ((uint64_t) y << 32) | x;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) r |= (uint32_t) ((pool >> (8 * ((s >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
  return r;
#endif
}

// bit 7 of every byte of the result is set iff that byte of x is zero (bytes do not interact)
CAB_HD uint32_t zero_bytes(uint32_t x) { return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u; }
// 0xff in every byte whose bit 7 is set in m (m has only bits 7, 15, 23, 31)
CAB_HD uint32_t spread(uint32_t m) { return (m >> 7) * 0xffu; }

// w0 w1 w2 = 12 record bytes {type, colour, flag} x 4 squares (little endian); returns the 4 int8 codes
CAB_HD uint32_t encode4(uint32_t w0, uint32_t w1, uint32_t w2) {
  // de-interleave: bytes 0 3 6 9 are the types, 1 4 7 10 the colours, 2 5 8 11 the flags
  const uint32_t T = byte_perm(byte_perm(w0, w1, 0x0630u), w2, 0x5210u);
  const uint32_t C = byte_perm(byte_perm(w0, w1, 0x0741u), w2, 0x6210u);
  const uint32_t F = byte_perm(byte_perm(w0, w1, 0x0052u), w2, 0x7410u);
  // type -> table index 0..7; a byte with any of bits 3..7 set is no piece type: index 7 (every table holds 0 there)
  const uint32_t junk = ((((T >> 3) & 0x1f1f1f1fu) + 0x7f7f7f7fu) & 0x80808080u) >> 7;
  const uint32_t t3 = (T & 0x07070707u) | (junk * 7u);
  const uint32_t x = t3 | (t3 >> 4);
  const uint32_t sel = byte_perm(x, 0u, 0x4420u);                       // nibbles t0 t1 t2 t3
  const uint32_t pos = byte_perm(0x04030100u, 0x00080706u, sel);        //  0  1  3  4  6  7  8  0
  const uint32_t neg = byte_perm(0xfcfdff00u, 0x00f8f9fau, sel);        //  0 -1 -3 -4 -6 -7 -8  0
  const uint32_t bon = byte_perm(0x01000100u, 0x00010000u, sel)         // king, rook, pawn: + 1 ...
                       & ((~zero_bytes(F) & 0x80808080u) >> 7);         // ... when the flag byte is not zero
  const uint32_t white = spread(zero_bytes(C ^ 0x01010101u)), black = spread(zero_bytes(C ^ 0x02020202u));
  // pos + bon <= 9 and neg - bon >= -9 where bon != 0: no carry or borrow leaves a byte
  return ((pos + bon) & white) | ((neg - bon) & black);
}

}  // namespace cab
