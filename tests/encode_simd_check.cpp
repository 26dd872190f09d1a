// Host check of csrc/encode_simd.cuh (the encoder kernel's byte-parallel rule) against the scalar rule of
// Piece::code()/0.4 (/root/reference/src/chess/piece.cpp:103-105,145-147,179-181,260-262): every (type, colour) byte pair,
// flags at both ends of the byte range and a stride between, in each of the four square positions, random neighbours.
// Built and run by tests/test_abi_host.py with g++ (no GPU).
#include "encode_simd.cuh"
#include <cstdio>
#include <cstdlib>
static int scalar(unsigned type, unsigned color, unsigned flag) {
  int k = 0;
  if (type >= 1 && type <= 6 && (color == 1 || color == 2)) {
    k = (int) ((0x876431u >> (4 * (type - 1))) & 0xfu);
    if (flag != 0 && (type == 1 || type == 3 || type == 6)) k += 1;
    if (color == 2) k = -k;
  }
  return k;
}
int main() {
  unsigned long bad = 0, n = 0;
  unsigned seed = 12345;
  for (unsigned t = 0; t < 256; ++t) for (unsigned c = 0; c < 256; ++c) for (unsigned f = 0; f < 256; f += (f < 4 || f > 250) ? 1 : 13) {
    unsigned char b[12];
    for (int i = 0; i < 12; ++i) { seed = seed * 1664525u + 1013904223u; b[i] = (seed >> 24) % ((seed >> 8) % 3 == 0 ? 256 : 9); }
    for (int s = 0; s < 4; ++s) {
      unsigned char bb[12];
      for (int i = 0; i < 12; ++i) bb[i] = b[i];
      bb[3 * s] = t; bb[3 * s + 1] = c; bb[3 * s + 2] = f;
      uint32_t w[3];
      for (int i = 0; i < 3; ++i) w[i] = bb[4 * i] | (bb[4 * i + 1] << 8) | (bb[4 * i + 2] << 16) | ((uint32_t) bb[4 * i + 3] << 24);
      const uint32_t r = cab::encode4(w[0], w[1], w[2]);
      for (int q = 0; q < 4; ++q) {
        const int want = scalar(bb[3 * q], bb[3 * q + 1], bb[3 * q + 2]);
        if ((int) (int8_t) (r >> (8 * q)) != want) { if (bad++ < 5) printf("bad q=%d t=%u c=%u f=%u got %d want %d\n", q, bb[3*q], bb[3*q+1], bb[3*q+2], (int)(int8_t)(r >> (8*q)), want); }
        ++n;
      }
    }
  }
  printf("checked %lu bad %lu\n", n, bad);
  return bad != 0;
}
