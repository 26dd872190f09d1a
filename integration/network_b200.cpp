// network_b200.cpp — drop-in replacement for the reference's src/mcts_network/network.cpp.
//
// It implements every member the reference's UNMODIFIED header src/mcts_network/network.h declares
// (network::Network, network::Optimizer, network::NetworkStorage, operator>> / operator<<) on top of the C ABI of
// include/chessai_b200.h, so src/player/player.cpp, src/mcts_network/{decider,tree}.cpp and src/main/** compile and
// link unchanged and the network runs on the B200. Build: compile this file INSTEAD of network.cpp with
// -I<reference>/src -I<this repo>/include and link libchessai_b200.so (integration/Makefile).
//
// Written from the header's contract, not from network.cpp. Where behaviour is specified by the reference
// implementation the line is cited (paths under /root/reference/src/mcts_network/).
//
// Device state: each Network owns one cab_net (looked up by object address; the header leaves no room for a new
// member). The header's host vectors (_connections) are kept as a lazily refreshed mirror for operator<<.
// There is no CPU fallback: without a GPU every compute member aborts through DEBUG_ASSERT like the reference does
// on its own fatal paths (CMakeLists.txt:46-49 always defines DEBUG).
#include "mcts_network/network.h"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <shared_mutex>
#include <unordered_map>

#include "chessai_b200.h"
#include "mcts_network/tree.h"
#include "util/string_util.h"

namespace {

struct DeviceSide {
  cab_net *handle = nullptr;
  bool host_stale = false;  // the device weights moved on (optimize / dropout); refresh before reading _connections
};
std::shared_mutex g_mu;
std::unordered_map<const network::Network *, DeviceSide> g_dev;

[[noreturn]] void die(const char *what) {
  std::fprintf(stderr, "chessai_b200: %s failed: %s\n", what, cab_last_error());
  DEBUG_ASSERT
  std::abort();
}
void ck(int rc, const char *what) {
  if (rc != CAB_OK) die(what);
}

int device_index() {
  const char *e = std::getenv("CHESSAI_B200_DEVICE");
  return e ? std::atoi(e) : 0;
}

DeviceSide &side_of(const network::Network *n) {
  std::shared_lock<std::shared_mutex> lk(g_mu);
  auto it = g_dev.find(n);
  if (it == g_dev.end()) {
    std::fprintf(stderr, "chessai_b200: network %p has no device state\n", (const void *) n);
    std::abort();
  }
  return it->second;
}

long weight_count(const std::vector<int> &d) {
  long c = 0;
  for (size_t l = 0; l + 1 < d.size(); ++l) c += (long) d[l] * d[l + 1];
  return c;
}

// Network::loadBoard's integer stage (network.cpp:119-123): Piece::code() is k * 0.4 exactly, k in [-9, 9]
void encode(game::Board *b, int8_t *codes) {
  std::vector<piece::Piece *> pcs = b->pieces();
  for (int i = 0; i < 64; ++i) codes[i] = (int8_t) std::lround(pcs[i]->code() * 2.5);
}

}  // namespace

namespace network {

// A friend of Network through the header's `friend class NetworkManager`: the flat <-> nested weight copies
struct Bridge : NetworkManager {
  static void upload(Network *n) {
    auto &w = get_connections(n);
    auto &d = get_dimensions(n);
    std::vector<double> flat;
    flat.reserve(weight_count(d));
    for (size_t l = 0; l + 1 < d.size(); ++l)
      for (int i = 0; i < d[l]; ++i)
        for (int j = 0; j < d[l + 1]; ++j) flat.push_back(w[l][i][j]);
    DeviceSide &s = side_of(n);
    ck(cab_net_set_weights(s.handle, flat.data(), (long) flat.size()), "cab_net_set_weights");
    s.host_stale = false;
  }
  static void download(Network *n) {
    DeviceSide &s = side_of(n);
    if (!s.host_stale) return;
    auto &w = get_connections(n);
    auto &d = get_dimensions(n);
    std::vector<double> flat(weight_count(d));
    ck(cab_net_get_weights(s.handle, flat.data(), (long) flat.size()), "cab_net_get_weights");
    long c = 0;
    for (size_t l = 0; l + 1 < d.size(); ++l)
      for (int i = 0; i < d[l]; ++i)
        for (int j = 0; j < d[l + 1]; ++j) w[l][i][j] = flat[c++];
    s.host_stale = false;
  }
  static std::vector<int> &dims(Network *n) { return get_dimensions(n); }
  static std::vector<std::vector<std::vector<double>>> &weights(Network *n) { return get_connections(n); }
};

// ---- Network ---------------------------------------------------------------------------------------------------
Network::Network() : Network({64, 144, 225, 100, 1}) {}

Network::Network(std::vector<int> dims) {
  // dims normalisation as the reference performs it (network.cpp:39-56); the C ABI applies the same rule
  if (dims.empty()) dims = {64, 144, 225, 100, 1};
  else {
    if (dims[0] != 64) dims.insert(dims.begin(), 64);
    if (dims.back() != 1) dims.push_back(1);
    if (dims.size() <= 2) dims = {64, 144, 225, 100, 1};
  }
  _num_layers = 0;
  updateInternals(dims);
}

void Network::updateInternals(const std::vector<int> &dims) {
  // (re)shape the host mirror and (re)create the device network. Unlike the reference (which appends to the old
  // vectors) a second call replaces the first, so dumps of any shape load correctly.
  _num_layers = (int) dims.size();
  _input_index = 0;
  _output_index = _num_layers - 1;
  _dimensions = dims;
  _max_width = 0;
  _neurons.clear(); _thetas.clear(); _omegas.clear(); _connections.clear(); _connection_changes.clear();
  for (int n = 0; n < _num_layers; ++n) {
    _neurons.emplace_back(dims[n], 0.0);
    _thetas.emplace_back(dims[n], 0.0);
    _omegas.emplace_back(dims[n], 0.0);
    _max_width = std::max(_max_width, dims[n]);
  }
  for (int n = 0; n + 1 < _num_layers; ++n) {
    _connections.emplace_back(dims[n], std::vector<double>(dims[n + 1], 0.0));
    _connection_changes.emplace_back(dims[n], std::vector<double>(dims[n + 1], 0.0));
  }
  std::unique_lock<std::shared_mutex> lk(g_mu);
  DeviceSide &s = g_dev[this];
  if (s.handle != nullptr) cab_net_destroy(s.handle);
  s.handle = nullptr;
  ck(cab_net_create(dims.data(), (int) dims.size(), device_index(), &s.handle), "cab_net_create");
  s.host_stale = false;
}

Network::~Network() {
  std::unique_lock<std::shared_mutex> lk(g_mu);
  auto it = g_dev.find(this);
  if (it != g_dev.end()) {
    cab_net_destroy(it->second.handle);
    g_dev.erase(it);
  }
}

double Network::predictPosition(game::Board *b) {
  // re-entrant and read-only like the reference's (tree.cpp:185 calls it from N threads on one object):
  // cab_eval_host takes a private stream + buffers per caller
  int8_t codes[64];
  encode(b, codes);
  double y = 0.0;
  ck(cab_eval_host(side_of(this).handle, codes, 1, &y), "cab_eval_host");
  return y;
}

void Network::loadBoard(game::Board *b, std::vector<double> &input) {
  int8_t codes[64];
  encode(b, codes);
  for (int i = 0; i < _dimensions[0]; ++i) input[i] = codes[i] * 0.4;
}

// propagate_for_training: the per-layer activations of the board last put in _neurons[0] (network.cpp:125-153).
// Only Optimizer used them in the reference; kept for completeness of the header's surface.
void Network::propagate_for_training() {
  std::vector<std::vector<double>> unused = _thetas;
  propagate_for_training(unused);
}
void Network::propagate_for_training(std::vector<std::vector<double>> &unbounded) {
  int8_t codes[64];
  for (int i = 0; i < 64; ++i) codes[i] = (int8_t) std::lround(_neurons[0][i] * 2.5);
  long row = 0;
  for (int n = 1; n < _num_layers; ++n) row += _dimensions[n];
  std::vector<double> acts(row), thetas(row);
  ck(cab_forward_thetas_host(side_of(this).handle, codes, 1, acts.data(), thetas.data()), "cab_forward_thetas_host");
  long c = 0;
  for (int n = 1; n < _num_layers; ++n)
    for (int j = 0; j < _dimensions[n]; ++j, ++c) {
      _neurons[n][j] = acts[c];
      unbounded[n][j] = thetas[c];   // the sum itself (network.cpp:150), exact also for saturated units
    }
}

Network *Network::randomNetwork() {
  auto *n = new Network();
  randomNetwork(n);
  return n;
}
Network *Network::randomNetwork(const std::vector<int> &dims) {
  auto *n = new Network(dims);
  randomNetwork(n);
  return n;
}
void Network::randomNetwork(Network *network) {
  // U(-1,1) from the engine's own generator, in n,i,j order: the same stream the reference consumes (network.cpp:167-173)
  auto &w = Bridge::weights(network);
  for (auto &layer : w)
    for (auto &row : layer)
      for (auto &x : row) x = random(-1.0, 1.0);
  Bridge::upload(network);
}

Network *Network::uniformNetwork() {
  auto *n = new Network();
  uniformNetwork(n);
  return n;
}
Network *Network::uniformNetwork(const std::vector<int> &dims) {
  auto *n = new Network(dims);
  uniformNetwork(n);
  return n;
}
void Network::uniformNetwork(Network *network) {
  for (auto &layer : Bridge::weights(network))
    for (auto &row : layer)
      for (auto &x : row) x = 0.0;
  ck(cab_net_zero(side_of(network).handle), "cab_net_zero");
}

Network *Network::clone() {
  // weights only, like the reference (network.cpp:195-205); the copy gets its own device network
  Bridge::download(this);
  auto *n = new Network(_dimensions);
  n->_connections = _connections;
  Bridge::upload(n);
  return n;
}
Network *Network::clone(Network *network) { return network->clone(); }

Network *Network::loadFromFile(const std::string &file_path) {
  Network *network = nullptr;
  std::ifstream in(file_path);
  if (in.is_open()) {
    in >> network;
    in.close();
  } else DEBUG_ASSERT
  return network;
}

std::istream &operator>>(std::istream &input, Network *&net) {
  // "<L>\n<d0> ... <dL-1>\n<weights in n,i,j order>" (network.cpp:386-407)
  int layers = 0;
  input >> layers;
  std::vector<int> dims(layers > 0 ? layers : 0);
  for (int n = 0; n < layers; ++n) input >> dims[n];
  if (net == nullptr) net = new Network(dims);
  else net->updateInternals(dims);
  for (auto &layer : net->_connections)
    for (auto &row : layer)
      for (auto &x : row) input >> x;
  Bridge::upload(net);
  return input;
}

std::ostream &operator<<(std::ostream &output, Network *&net) {
  // byte-exact network_dump writer (network.cpp:409-426): every weight as string::from_double followed by a space
  if (net == nullptr) { DEBUG_ASSERT return output; }
  Bridge::download(net);
  output << net->_num_layers << std::endl;
  for (int n = 0; n + 1 < net->_num_layers; ++n) output << net->_dimensions[n] << " ";
  output << net->_dimensions[net->_num_layers - 1] << std::endl;
  for (auto &layer : net->_connections)
    for (auto &row : layer)
      for (auto &x : row) output << string::from_double(x) << " ";
  return output;
}

// ---- Optimizer -------------------------------------------------------------------------------------------------
void Optimizer::optimize(Network *network, game::Board *input, const double TARGET_OUTPUT, double &lambda,
                         const double MAX_LEARNING_CONSTANT, const double LEARNING_RATE) {
  // forward, backprop with in-place update, re-forward, accept (lambda *= rate) or reject (lambda /= rate, rollback):
  // network.cpp:223-279, executed on the device in the reference's operation order
  int8_t codes[64];
  encode(input, codes);
  DeviceSide &s = side_of(network);
  int accepted = 0;
  ck(cab_train_case(s.handle, codes, TARGET_OUTPUT, &lambda, MAX_LEARNING_CONSTANT, LEARNING_RATE, &accepted, nullptr, nullptr),
     "cab_train_case");
  s.host_stale = true;
}

void Optimizer::dropout(Network *network, double dropout_rate) {
  // "weak dilution" (network.cpp:281-292). The engine's sequential mt19937 cannot drive a parallel kernel; one draw
  // from it seeds the Philox stream instead, so runs stay reproducible from the same generator.
  DeviceSide &s = side_of(network);
  static uint64_t offset = 0;
  const uint64_t seed = (uint64_t) (math::random() * 9007199254740992.0);
  ck(cab_dropout(s.handle, dropout_rate, seed, offset++, nullptr), "cab_dropout");
  s.host_stale = true;
}

// ---- NetworkStorage ----------------------------------------------------------------------------------------------
Network *NetworkStorage::_current_network = nullptr;
long NetworkStorage::_network_count = 0;
bool NetworkStorage::SAVE_NETWORKS = false;
std::function<void(game::Board *, double)> NetworkStorage::_network_training_case;

Network *NetworkStorage::current_network() {
  if (_current_network == nullptr) DEBUG_ASSERT
  return _current_network;
}

Network *NetworkStorage::initialize() {
  if (_current_network != nullptr) { DEBUG_ASSERT delete _current_network; }
  _current_network = Network::randomNetwork();
  saveLatestNetwork();
  return _current_network;
}
Network *NetworkStorage::initialize(const std::vector<int> &dims) {
  if (_current_network != nullptr) { DEBUG_ASSERT delete _current_network; }
  _current_network = Network::randomNetwork(dims);
  saveLatestNetwork();
  return _current_network;
}
Network *NetworkStorage::initialize(const std::string &file_path) {
  if (_current_network != nullptr) { DEBUG_ASSERT delete _current_network; }
  _current_network = Network::loadFromFile(file_path);
  saveLatestNetwork();
  return _current_network;
}

void NetworkStorage::saveNetwork() {
  // latest.txt -> gen-<N>.txt, then a fresh latest.txt (network.cpp:345-355)
  if (!SAVE_NETWORKS || _current_network == nullptr) return;
  if (_network_count >= 0) {
    const std::string past = NETWORK_DUMP_DIRECTORY + "gen-" + std::to_string(_network_count) + ".txt";
    std::rename(LATEST_NETWORK_FILE_PATH.c_str(), past.c_str());
  }
  ++_network_count;
  saveLatestNetwork();
}

void NetworkStorage::saveLatestNetwork() {
  if (!SAVE_NETWORKS) return;
  std::ofstream out(LATEST_NETWORK_FILE_PATH);
  out << _current_network;
}

void NetworkStorage::flushStorage() {
  saveNetwork();
  delete _current_network;
  _current_network = nullptr;
}

void NetworkStorage::saveBoard(const game::Board *board, const tree::Node *node) {
  if (!_network_training_case) return;
  game::Board *copy = board->clone();
  _network_training_case(copy, node->value());
  delete copy;
}
void NetworkStorage::setTestCaseSelector(const std::function<void(game::Board *, double)> &selector) {
  _network_training_case = selector;
}

}  // namespace network
