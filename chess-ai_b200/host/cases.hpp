// cases.hpp — training cases: the reference's target rule and text format, plus packed binary shards.
//
// Reference (all under /root/reference/src/main):
//   initialization.cpp:221-227  SIMULATION_BOARD_SAVE_PROCEDURE: with probability `ratio` keep the board reached after
//                               the move, with  b = clamp(2*v - 1, -1, 1),  v = value() of the searched root
//                               (mcts_network/network.cpp:372-378, player/player.cpp:455-463)
//   network/make_cases.cpp:36-42   get_overall_result: target = 10 * (4*b + game_result) / 5, game_result in {-1,0,1}
//   network/make_cases.cpp:113-129 case file = board text ("8 8", 64 lines "<i> - <type> <colour> <flag>") + target line
// Packed shard: a headerless array of 72-byte records { int8 codes[64]; double target; } (little endian), the layout
// cab_train_batch_host consumes after a split into codes[] and targets[]. Shards concatenate with `cat`.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace cases {

constexpr int kRecordBytes = 72;

struct Case {
  int8_t codes[64];
  double target;
};
static_assert(sizeof(Case) == kRecordBytes, "packed case record is 72 bytes");

inline double bounded_mcts_result(double node_value) {    // initialization.cpp:223-224
  double b = 2.0 * node_value - 1.0;
  if (b < -1.0) b = -1.0;
  if (b > 1.0) b = 1.0;
  return b;
}

inline double overall_result(double mcts_result, double game_result, double scale = 10.0) {   // make_cases.cpp:36-42
  const double mcts_weight = 4.0, game_result_weight = 1.0, weights_sum = mcts_weight + game_result_weight;
  return scale * (mcts_weight * mcts_result + game_result_weight * game_result) / weights_sum;
}

inline bool append_shard(const std::string &path, const std::vector<Case> &cs) {
  FILE *f = std::fopen(path.c_str(), "ab");
  if (!f) return false;
  const size_t n = std::fwrite(cs.data(), kRecordBytes, cs.size(), f);
  return std::fclose(f) == 0 && n == cs.size();
}

inline bool read_shard(const std::string &path, std::vector<Case> &out) {
  FILE *f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  const long bytes = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  if (bytes < 0 || bytes % kRecordBytes != 0) { std::fclose(f); return false; }
  const size_t at = out.size(), n = (size_t) bytes / kRecordBytes;
  out.resize(at + n);
  const size_t got = std::fread(out.data() + at, kRecordBytes, n, f);
  std::fclose(f);
  return got == n;
}

// code k in [-9,9] -> "<type> <colour> <flag>" (chess/piece.cpp:300-375; the flag exists for king / rook / pawn only)
inline const char *piece_words(int k) {
  static const char *const kWhite[10] = {"none none .", "king white false", "king white true", "queen white .",
                                         "rook white false", "rook white true", "knight white .", "bishop white .",
                                         "pawn white false", "pawn white true"};
  static const char *const kBlack[10] = {"none none .", "king black false", "king black true", "queen black .",
                                         "rook black false", "rook black true", "knight black .", "bishop black .",
                                         "pawn black false", "pawn black true"};
  if (k < -9 || k > 9) return nullptr;
  return k >= 0 ? kWhite[k] : kBlack[-k];
}

// the text of one case file, byte for byte what network::save_training_case writes
inline std::string case_text(const Case &c) {
  std::string s = "8 8\n";
  char line[64];
  for (int i = 0; i < 64; ++i) {
    std::snprintf(line, sizeof line, "%d - %s\n", i, piece_words(c.codes[i]));
    s += line;
  }
  std::snprintf(line, sizeof line, "%.17e\n", c.target);   // string::from_double (util/string_util.cpp:39-43)
  s += line;
  return s;
}

inline bool parse_case_text(const std::string &text, Case &c) {
  const char *p = text.c_str();
  int w = 0, h = 0, used = 0;
  if (std::sscanf(p, "%d %d%n", &w, &h, &used) != 2 || w != 8 || h != 8) return false;
  p += used;
  for (int i = 0; i < 64; ++i) {
    int idx = 0;
    char type[16], colour[16], flag[16];
    if (std::sscanf(p, "%d - %15s %15s %15s%n", &idx, type, colour, flag, &used) != 4 || idx != i) return false;
    p += used;
    int base;
    bool has_flag = false;
    if (!std::strcmp(type, "none")) base = 0;
    else if (!std::strcmp(type, "king")) { base = 1; has_flag = true; }
    else if (!std::strcmp(type, "queen")) base = 3;
    else if (!std::strcmp(type, "rook")) { base = 4; has_flag = true; }
    else if (!std::strcmp(type, "knight")) base = 6;
    else if (!std::strcmp(type, "bishop")) base = 7;
    else if (!std::strcmp(type, "pawn")) { base = 8; has_flag = true; }
    else return false;
    if (has_flag && !std::strcmp(flag, "true")) ++base;
    const int sign = !std::strcmp(colour, "white") ? 1 : !std::strcmp(colour, "black") ? -1 : 0;
    c.codes[i] = (int8_t) (sign * base);
  }
  return std::sscanf(p, "%lf", &c.target) == 1;
}

}  // namespace cases
