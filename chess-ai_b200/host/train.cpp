// train.cpp — the trainer loop over packed case shards, with checkpoint rotation (SURVEY §8f N3 + N4).
//
// Reference: network::train::train_network (/root/reference/src/main/network/train.cpp:39-68): lambda0 = 2.0; draw a
// case uniformly, Optimizer::optimize, when lambda leaves [1e-4, 10] reset it to 2.0 and dropout; every
// GAMES_PER_NETWORK_SAVE cases NetworkStorage::saveNetwork (network.cpp:345-364): rename latest.txt -> gen-<N>.txt,
// write the new latest.txt.
//   mode "exact": that loop verbatim, one case per step, executed on the device by cab_train_sequence
//   mode "batch": the minibatch extension as resident epochs (cab_train_epoch_host), same lambda / dropout rule per step, on the device
// Beyond the reference: lambda, the step count and the generation number persist in <out_dir>/trainer_state.txt, so a
// restarted trainer continues instead of starting again at lambda0 and overwriting gen-0 (the reference loses both).
//
// Data-parallel (batch mode): one process per GPU; rank r trains on cases r, r + world, r + 2 world, ... of the shard,
// the library all-reduces the gradient (cab_dp_init: NCCL), every rank takes the identical decision and applies the
// identical update; rank 0 alone writes checkpoints. Rendezvous without MPI: rank 0 writes the 128-byte NCCL id to
// <id_file>, the others poll for it.
//
//   train <net_in.txt> <shard> <out_dir> <exact|batch> <steps> <batch> [device] [seed] [save_every] [rank world id_file]
// Prints one JSON object (rank 0). No CPU fallback.
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../include/chessai_b200.h"
#include "cases.hpp"

namespace {

struct TrainerState {
  double lambda = 2.0;   // Optimizer::DEFAULT_INITIAL_LAMBDA, network.h:153
  long steps = 0, resets = 0, generation = 0;
};

bool load_state(const std::string &path, TrainerState &s) {
  FILE *f = std::fopen(path.c_str(), "r");
  if (!f) return false;
  const int n = std::fscanf(f, "%lf %ld %ld %ld", &s.lambda, &s.steps, &s.resets, &s.generation);
  std::fclose(f);
  return n == 4;
}

bool save_state(const std::string &path, const TrainerState &s) {
  FILE *f = std::fopen(path.c_str(), "w");
  if (!f) return false;
  std::fprintf(f, "%.17e %ld %ld %ld\n", s.lambda, s.steps, s.resets, s.generation);
  return std::fclose(f) == 0;
}

// NetworkStorage::saveNetwork: latest.txt becomes gen-<N>.txt, the current weights become latest.txt
int rotate(cab_net *net, const std::string &dir, TrainerState &s) {
  const std::string latest = dir + "/latest.txt";
  FILE *probe = std::fopen(latest.c_str(), "r");
  if (probe) {
    std::fclose(probe);
    const std::string past = dir + "/gen-" + std::to_string(s.generation++) + ".txt";
    if (std::rename(latest.c_str(), past.c_str()) != 0) return CAB_EIO;
  }
  const int rc = cab_net_dump_text(net, latest.c_str());
  if (rc != CAB_OK) return rc;
  return save_state(dir + "/trainer_state.txt", s) ? CAB_OK : CAB_EIO;
}

#define CK(call) do { if ((call) != CAB_OK) { std::fprintf(stderr, "%s: %s\n", #call, cab_last_error()); return 1; } } while (0)

}  // namespace

int main(int argc, char **argv) {
  if (argc < 7) {
    std::fprintf(stderr, "usage: %s net_in.txt shard out_dir exact|batch steps batch [device] [seed] [save_every]\n", argv[0]);
    return 2;
  }
  const std::string net_in = argv[1], shard = argv[2], out_dir = argv[3], mode = argv[4];
  const long steps = std::atol(argv[5]);
  const long batch = mode == "exact" ? 1 : std::atol(argv[6]);
  const int device = argc > 7 ? std::atoi(argv[7]) : 0;
  const unsigned long long seed = argc > 8 ? std::strtoull(argv[8], nullptr, 10) : 5489;
  const long save_every = argc > 9 ? std::atol(argv[9]) : 20;   // GAMES_PER_NETWORK_SAVE default, initialization.h:56
  const int rank = argc > 12 ? std::atoi(argv[10]) : 0, world = argc > 12 ? std::atoi(argv[11]) : 1;
  const std::string id_file = argc > 12 ? argv[12] : "";
  if (world > 1 && mode != "batch") { std::fprintf(stderr, "data-parallel training needs batch mode\n"); return 2; }
  if ((mode != "exact" && mode != "batch") || steps <= 0 || batch <= 0 || save_every <= 0) { std::fprintf(stderr, "bad arguments\n"); return 2; }

  std::vector<cases::Case> cs;
  if (!cases::read_shard(shard, cs) || cs.empty()) { std::fprintf(stderr, "cannot read shard %s\n", shard.c_str()); return 1; }
  if (world > 1) {                              // this rank's stride of the shard
    std::vector<cases::Case> mine;
    for (size_t i = (size_t) rank; i < cs.size(); i += (size_t) world) mine.push_back(cs[i]);
    cs.swap(mine);
    if (cs.empty()) { std::fprintf(stderr, "shard smaller than the world size\n"); return 1; }
  }
  TrainerState st;
  const bool resumed = load_state(out_dir + "/trainer_state.txt", st);
  cab_net *net = nullptr;
  CK(cab_net_load_text(resumed ? (out_dir + "/latest.txt").c_str() : net_in.c_str(), device, &net));
  if (world > 1) {
    unsigned char id[128];
    if (rank == 0) {
      CK(cab_dp_unique_id(id));
      const std::string tmp = id_file + ".tmp";
      FILE *f = std::fopen(tmp.c_str(), "wb");
      if (!f || std::fwrite(id, 1, 128, f) != 128 || std::fclose(f) != 0 || std::rename(tmp.c_str(), id_file.c_str()) != 0) {
        std::fprintf(stderr, "cannot write %s\n", id_file.c_str());
        return 1;
      }
    } else {
      FILE *f = nullptr;
      for (int tries = 0; tries < 600 && !(f = std::fopen(id_file.c_str(), "rb")); ++tries) usleep(100000);
      if (!f || std::fread(id, 1, 128, f) != 128) { std::fprintf(stderr, "cannot read %s\n", id_file.c_str()); return 1; }
      std::fclose(f);
    }
    CK(cab_dp_init(net, id, rank, world));
  }
  const bool writer = rank == 0;                // only rank 0 touches the checkpoint directory

  std::mt19937_64 rng(seed + (unsigned long long) st.steps);
  std::vector<int8_t> codes((size_t) std::max(batch, save_every) * 64);
  std::vector<double> targets((size_t) std::max(batch, save_every));
  auto draw = [&](long n) {                     // train.cpp:49: uniform draw with replacement
    for (long i = 0; i < n; ++i) {
      const cases::Case &c = cs[rng() % cs.size()];
      std::memcpy(&codes[(size_t) i * 64], c.codes, 64);
      targets[(size_t) i] = c.target;
    }
  };

  long accepted = 0, decisions = 0, samples = 0, resets0 = st.resets, gen0 = st.generation;
  double first_err = -1.0, last_err = -1.0;
  const auto t0 = std::chrono::steady_clock::now();
  if (mode == "exact") {
    std::vector<uint8_t> acc((size_t) save_every);
    for (long done = 0; done < steps;) {
      const long n = std::min(save_every, steps - done);
      draw(n);
      int resets = 0;
      CK(cab_train_sequence(net, codes.data(), targets.data(), n, &st.lambda, seed + (unsigned long long) st.resets, acc.data(), &resets));
      for (long i = 0; i < n; ++i) accepted += acc[(size_t) i];
      decisions += n; samples += n; done += n;
      st.steps += n; st.resets += resets;
      if (writer && n == save_every) CK(rotate(net, out_dir, st));
    }
  } else {
    // one save interval at a time: its minibatches are drawn on the host, uploaded once and trained as a RESIDENT epoch
    // (cab_train_epoch_host: lambda, accept / reject, rollback and the reset + dropout rule of train.cpp:54-57 on the device)
    codes.resize((size_t) save_every * batch * 64);
    targets.resize((size_t) save_every * batch);
    for (long done = 0; done < steps;) {
      const long n_steps = std::min(save_every - st.steps % save_every, steps - done);
      draw(n_steps * batch);
      cab_train_stats ts;
      CK(cab_train_epoch_host(net, codes.data(), targets.data(), n_steps * batch, batch, &st.lambda, 15.0, 1.1, 1,
                              seed + (unsigned long long) st.resets, &ts, nullptr));
      if (first_err < 0.0) first_err = ts.first_err;
      last_err = ts.last_accepted ? ts.last_new_err : ts.last_err;
      accepted += ts.accepted; decisions += ts.steps; samples += n_steps * batch; done += n_steps;
      st.steps += n_steps; st.resets += ts.resets;
      if (writer && st.steps % save_every == 0) CK(rotate(net, out_dir, st));
    }
  }
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (writer) CK(rotate(net, out_dir, st));     // NetworkStorage::flushStorage, network.cpp:366-370
  long launches = 0;
  cab_net_launch_count(net, &launches);
  if (world > 1) CK(cab_dp_shutdown(net));
  if (writer) std::printf("{\"config\": \"trainer over packed case shards\", \"mode\": \"%s\", \"cases_in_shard\": %zu, \"steps\": %ld, \"batch\": %ld, "
              "\"resumed\": %s, \"seconds\": %.3f, \"samples\": %ld, \"samples_per_s\": %.1f, \"accepted\": %ld, \"decisions\": %ld, "
              "\"first_err\": %.9e, \"last_err\": %.9e, \"lambda\": %.17e, \"resets\": %ld, \"generations_written\": %ld, "
              "\"total_steps\": %ld, \"gpu_launches\": %ld, \"world\": %d}\n",
              mode.c_str(), cs.size(), steps, batch, resumed ? "true" : "false", secs, samples * world, samples * world / secs, accepted, decisions,
              first_err, last_err, st.lambda, st.resets - resets0, st.generation - gen0, st.steps, launches, world);
  cab_net_destroy(net);
  return 0;
}
