"""chess-ai_b200 — B200-native MLP evaluator / trainer for utk003/Chess-AI (Python host mirror).

The product is libchessai_b200.so (hand-written sm_100a CUDA behind the C ABI of include/chessai_b200.h).
This module is a thin ctypes mirror of the reference's class surface for that path —
network::Network / network::Optimizer / network::NetworkStorage (/root/reference/src/mcts_network/network.h:46-183)
— used by the tests and bench.py. There is NO CPU fallback: if the CUDA library is missing or no GPU is
present, compute calls raise CabError.

Import with importlib (the directory name carries a hyphen):
    cab = importlib.import_module("chess-ai_b200")
"""
import ctypes as C
import os

import numpy as np

from . import build as _build
from .synth import synthetic_boards  # noqa: F401  (re-export)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = _build.LIB
HEADER = os.path.join(os.path.dirname(HERE), "include", "chessai_b200.h")

CAB_OK, CAB_EINVAL, CAB_ENODEVICE, CAB_ECUDA, CAB_EIO, CAB_ENCCL, CAB_ESTATE = range(7)


class CabError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("chessai_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None
_vp = C.c_void_p
_i8p = np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_long)

# name -> (argtypes). Every function returns int except cab_version / cab_last_error.
_SIGS = {
    "cab_device_count": [_ip],
    "cab_net_create": [_i32p, C.c_int, C.c_int, C.POINTER(_vp)],
    "cab_net_destroy": [_vp],
    "cab_net_clone": [_vp, C.POINTER(_vp)],
    "cab_net_num_layers": [_vp, _ip],
    "cab_net_dims": [_vp, _i32p],
    "cab_net_num_weights": [_vp, _lp],
    "cab_net_device": [_vp, _ip],
    "cab_net_set_weights": [_vp, _f64p, C.c_long],
    "cab_net_get_weights": [_vp, _f64p, C.c_long],
    "cab_net_randomize": [_vp, C.c_uint64, C.c_uint64],
    "cab_net_zero": [_vp],
    "cab_net_load_text": [C.c_char_p, C.c_int, C.POINTER(_vp)],
    "cab_net_dump_text": [_vp, C.c_char_p],
    "cab_encode_host": [C.c_int, _u8p, C.c_long, _i8p],
    "cab_encode_dev": [_vp, C.c_long, _vp, _vp],
    "cab_eval_host": [_vp, _vp, C.c_long, _vp],
    "cab_eval_dev": [_vp, _vp, C.c_long, _vp, _vp],
    "cab_eval_dev_batches": [_vp, _vp, C.c_long, C.c_long, _vp, C.POINTER(_vp), C.c_int],
    "cab_leafq_create": [_vp, C.c_long, C.c_long, C.c_int, C.POINTER(_vp)],
    "cab_leafq_destroy": [_vp],
    "cab_leafq_push": [_vp, _i8p, C.c_long, _lp],
    "cab_leafq_pending": [_vp, _lp],
    "cab_leafq_set_precision": [_vp, C.c_int],
    "cab_prediction_host": [_vp, _i8p, C.c_int, C.c_int, _dp, C.POINTER(C.c_int), _f64p, C.c_int, _ip],
    "cab_best_move_host": [_vp, _i8p, C.c_int, C.c_int, C.c_double, _ip, _ip, _dp, _ip, _ip],
    "cab_leafq_flush": [_vp, _lp],
    "cab_leafq_flush_async": [_vp, _lp],
    "cab_leafq_wait": [_vp],
    "cab_leafq_results": [_vp, C.POINTER(_dp)],
    "cab_leafq_stats": [_vp, _lp, _lp, _lp],
    "cab_prior_softmax_host": [C.c_int, _f64p, _i32p, C.c_int, _f64p, _f64p],
    "cab_eval_bf16_dev": [_vp, _vp, C.c_long, _vp, _vp],
    "cab_eval_bf16_host": [_vp, _i8p, C.c_long, _f64p],
    "cab_eval_records_host": [_vp, _u8p, C.c_long, _f64p],
    "cab_forward_full_host": [_vp, _i8p, C.c_long, _f64p],
    "cab_forward_thetas_host": [_vp, _i8p, C.c_long, _f64p, _f64p],
    "cab_train_case": [_vp, _i8p, C.c_double, _dp, C.c_double, C.c_double, _ip, _dp, _dp],
    "cab_train_sequence": [_vp, _i8p, _f64p, C.c_long, _dp, C.c_uint64, _vp, _ip],
    "cab_train_batch_host": [_vp, _i8p, _f64p, C.c_long, _dp, C.c_double, C.c_double, _ip, _dp, _dp],
    "cab_train_batch_dev": [_vp, _vp, _vp, C.c_long, _dp, C.c_double, C.c_double, _ip, _dp, _dp, _vp],
    "cab_train_epoch_dev": [_vp, _vp, _vp, C.c_long, C.c_long, _dp, C.c_double, C.c_double, C.c_int, C.c_uint64, _vp, _vp, _vp],
    "cab_train_epoch_host": [_vp, _i8p, _f64p, C.c_long, C.c_long, _dp, C.c_double, C.c_double, C.c_int, C.c_uint64, _vp, _vp],
    "cab_train_epoch_bf16_dev": [_vp, _vp, _vp, C.c_long, C.c_long, _dp, C.c_double, C.c_double, C.c_int, C.c_uint64, _vp, _vp, _vp],
    "cab_dropout": [_vp, C.c_double, C.c_uint64, C.c_uint64, _lp],
    "cab_dp_unique_id": [_vp],
    "cab_dp_init": [_vp, _vp, C.c_int, C.c_int],
    "cab_dp_shutdown": [_vp],
    "cab_dp_peer_export": [_vp, _vp],
    "cab_dp_peer_attach": [_vp, _vp, C.c_int, C.c_int],
    "cab_grad_buffer": [_vp, C.POINTER(_vp), _lp],
    "cab_eval_f16_dev": [_vp, _vp, C.c_long, _vp, _vp],
    "cab_eval_f16_host": [_vp, _i8p, C.c_long, _f64p],
    "cab_eval_f16_dev_batches": [_vp, _vp, C.c_long, C.c_long, _vp, C.POINTER(_vp), C.c_int],
    "cab_train_batch_bf16_host": [_vp, _i8p, _f64p, C.c_long, _dp, C.c_double, C.c_double, _ip, _dp, _dp],
    "cab_train_batch_bf16_dev": [_vp, _vp, _vp, C.c_long, _dp, C.c_double, C.c_double, _ip, _dp, _dp, _vp],
    "cab_grad_bf16_dev": [_vp, _vp, _vp, C.c_long, _vp],
    "cab_mcts_create": [_vp, _vp, C.POINTER(_vp)],
    "cab_mcts_destroy": [_vp],
    "cab_mcts_set_positions": [_vp, _i8p, _i8p, _vp],
    "cab_mcts_run": [_vp, C.c_int, C.c_long, _vp],
    "cab_mcts_root_children": [_vp, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _f64p, _i32p, _f64p, _ip, _ip, _dp],
    "cab_mcts_game_state": [_vp, C.c_int, _i8p, _ip, _ip, _ip, _ip],
    "cab_mcts_cases": [_vp, _vp, _vp, C.c_long, _lp],
    "cab_movegen_host": [C.c_int, _i8p, _i8p, C.c_int, C.c_int, C.c_int, _i32p, _u8p, _vp],
    "cab_measure_pipe_peak": [C.c_int, C.c_int, _dp],
    "cab_net_launch_count": [_vp, _lp],
}
ABI_SYMBOLS = ["cab_version", "cab_last_error"] + sorted(_SIGS)


def lib(build_if_missing=True):
    """The loaded C-ABI library. Fails loudly when it cannot be had: there is no other implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise CabError(CAB_ENODEVICE, "libchessai_b200.so is not built (run __graft_entry__.build())")
        _build.build()
    L = C.CDLL(LIB_PATH)
    L.cab_version.restype = C.c_int
    L.cab_last_error.restype = C.c_char_p
    for name, args in _SIGS.items():
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = args
    _lib = L
    return L


def _ck(rc):
    if rc != CAB_OK:
        raise CabError(rc, lib().cab_last_error().decode("utf-8", "replace"))


def device_count():
    n = C.c_int(0)
    rc = lib().cab_device_count(C.byref(n))
    return n.value if rc == CAB_OK else 0


def measure_pipe_peak(kind="dmma", device=0):
    """TFLOP/s of the fp64 tensor pipe ('dmma'), fp64 FMA pipe ('dfma') or fp32 FMA pipe ('ffma'), best of 10."""
    t = C.c_double(0)
    _ck(lib().cab_measure_pipe_peak(device, {"dmma": 0, "dfma": 1, "ffma": 2}[kind], C.byref(t)))
    return t.value


# ---- boards ------------------------------------------------------------------------------------------------------
PIECE_TYPES = ["none", "king", "queen", "rook", "knight", "bishop", "pawn"]
PIECE_COLORS = ["none", "white", "black"]


class Board:
    """64 piece records (type, colour, flag) — the slice of game::Board the network sees (chess/game.h:48-127).

    Text format = the reference's board / training-case files (chess/game.cpp:343-379, chess/piece.cpp:288-375):
    "8 8", then 64 lines "<idx> - <type> <colour> <flag>", optionally followed by the target value
    (main/network/make_cases.cpp:113-129).
    """

    def __init__(self, records=None):
        self.records = np.zeros((64, 3), dtype=np.uint8) if records is None else np.ascontiguousarray(records, dtype=np.uint8).reshape(64, 3)

    @classmethod
    def from_codes(cls, codes):
        t_of = [0, 1, 1, 2, 3, 3, 4, 5, 6, 6]
        f_of = [0, 0, 1, 0, 0, 1, 0, 0, 0, 1]
        rec = np.zeros((64, 3), dtype=np.uint8)
        for i, k in enumerate(np.asarray(codes).reshape(64)):
            a = abs(int(k))
            rec[i] = (t_of[a], 0 if a == 0 else (1 if k > 0 else 2), f_of[a])
        return cls(rec)

    @classmethod
    def from_text(cls, text):
        tok = text.split()
        if len(tok) < 2 + 64 * 5 or int(tok[0]) * int(tok[1]) != 64:
            raise CabError(CAB_EIO, "malformed board text")
        rec = np.zeros((64, 3), dtype=np.uint8)
        p = 2
        for i in range(64):
            if int(tok[p]) != i:
                raise CabError(CAB_EIO, "malformed board text: index %s at square %d" % (tok[p], i))
            t, c, m = tok[p + 2], tok[p + 3], tok[p + 4]
            rec[i] = (PIECE_TYPES.index(t) if t in PIECE_TYPES else 0, PIECE_COLORS.index(c) if c in PIECE_COLORS else 0, 1 if m == "true" else 0)
            p += 5
        target = float(tok[p]) if p < len(tok) else None
        return cls(rec), target

    @classmethod
    def load_training_case(cls, path):
        """network::load_training_case (make_cases.cpp:123-129) -> (Board, target)."""
        with open(path, "r") as f:
            return cls.from_text(f.read())

    def to_text(self, target=None):
        lines = ["8 8"]
        for i in range(64):
            t, c, m = (int(x) for x in self.records[i])
            flag = ("true" if m else "false") if t in (1, 3, 6) else "."
            lines.append("%d - %s %s %s" % (i, PIECE_TYPES[t], PIECE_COLORS[c], flag))
        s = "\n".join(lines) + "\n"
        if target is not None:
            s += "%.17e\n" % target
        return s

    def codes(self, device=0):
        """Network::loadBoard's integer stage, computed by the CUDA encoder kernel."""
        return encode_records(self.records[None], device)[0]


def encode_records(records, device=0):
    r = np.ascontiguousarray(records, dtype=np.uint8).reshape(-1, 64, 3)
    out = np.zeros((len(r), 64), dtype=np.int8)
    _ck(lib().cab_encode_host(device, r.reshape(-1), len(r), out.reshape(-1)))
    return out


def _as_codes(codes):
    a = np.ascontiguousarray(codes, dtype=np.int8)
    if a.shape[-1] != 64:
        raise CabError(CAB_EINVAL, "codes must have 64 entries per board")
    return a.reshape(-1, 64)


# ---- network::Network -----------------------------------------------------------------------------------------------
class Network:
    """Mirror of network::Network (network.h:46-107). Holds a cab_net resident on one GPU."""

    DEFAULT_DIMS = [64, 144, 225, 100, 1]

    def __init__(self, dims=None, device=0, _handle=None):
        self._h = _vp(None)
        if _handle is not None:
            self._h = _handle
            return
        d = np.ascontiguousarray(dims if dims is not None else self.DEFAULT_DIMS, dtype=np.int32)
        _ck(lib().cab_net_create(d, len(d), device, C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value and _lib is not None:
            _lib.cab_net_destroy(h)
            self._h = _vp(None)

    # -- construction helpers (network.cpp:155-220)
    @classmethod
    def randomNetwork(cls, dims=None, device=0, seed=5489, stream=0):
        n = cls(dims, device)
        _ck(lib().cab_net_randomize(n._h, seed, stream))
        return n

    @classmethod
    def uniformNetwork(cls, dims=None, device=0):
        return cls(dims, device)

    @classmethod
    def loadFromFile(cls, path, device=0):
        h = _vp(None)
        _ck(lib().cab_net_load_text(os.fsencode(path), device, C.byref(h)))
        return cls(_handle=h)

    @classmethod
    def from_weights(cls, dims, weights, device=0):
        n = cls(dims, device)
        n.set_weights(weights)
        return n

    def clone(self):
        h = _vp(None)
        _ck(lib().cab_net_clone(self._h, C.byref(h)))
        return Network(_handle=h)

    def dump(self, path):
        """operator<< (network.cpp:409-426): byte-exact network_dump text."""
        _ck(lib().cab_net_dump_text(self._h, os.fsencode(path)))

    # -- introspection
    @property
    def dims(self):
        n = C.c_int(0)
        _ck(lib().cab_net_num_layers(self._h, C.byref(n)))
        out = np.zeros(n.value, dtype=np.int32)
        _ck(lib().cab_net_dims(self._h, out))
        return [int(x) for x in out]

    @property
    def num_weights(self):
        n = C.c_long(0)
        _ck(lib().cab_net_num_weights(self._h, C.byref(n)))
        return n.value

    @property
    def launch_count(self):
        n = C.c_long(0)
        _ck(lib().cab_net_launch_count(self._h, C.byref(n)))
        return n.value

    def get_weights(self):
        out = np.zeros(self.num_weights, dtype=np.float64)
        _ck(lib().cab_net_get_weights(self._h, out, out.size))
        return out

    def set_weights(self, w):
        w = np.ascontiguousarray(w, dtype=np.float64).reshape(-1)
        _ck(lib().cab_net_set_weights(self._h, w, w.size))

    # -- evaluation (Network::predictPosition, network.cpp:92-117)
    def predictPosition(self, board):
        """One board (Board, or 64 int8 codes) -> evaluation in (-4, 4)."""
        if isinstance(board, Board):
            out = np.zeros(1, dtype=np.float64)
            _ck(lib().cab_eval_records_host(self._h, board.records.reshape(-1), 1, out))
            return float(out[0])
        return float(self.predict_batch(np.asarray(board).reshape(1, 64))[0])

    def predict_batch(self, codes, out=None):
        """Batched predictPosition over host int8 codes [B, 64] (H2D + kernel + D2H through the C ABI)."""
        c = _as_codes(codes)
        if out is None:
            out = np.zeros(len(c), dtype=np.float64)
        _ck(lib().cab_eval_host(self._h, c.ctypes.data, len(c), out.ctypes.data))
        return out

    def predict_records(self, records):
        r = np.ascontiguousarray(records, dtype=np.uint8).reshape(-1, 64, 3)
        out = np.zeros(len(r), dtype=np.float64)
        _ck(lib().cab_eval_records_host(self._h, r.reshape(-1), len(r), out))
        return out

    def eval_host_ptr(self, codes_ptr, n, out_ptr):
        """Raw-pointer form of predict_batch (pinned host buffers owned by the caller)."""
        _ck(lib().cab_eval_host(self._h, codes_ptr, n, out_ptr))

    def eval_dev_ptr(self, codes_dev_ptr, n, out_dev_ptr, stream=0):
        """Asynchronous evaluation on device pointers (codes int8 [n,64], out fp64 [n]) on a CUDA stream."""
        _ck(lib().cab_eval_dev(self._h, codes_dev_ptr, n, out_dev_ptr, stream))

    def eval_dev_batches(self, codes_dev_ptr, n, batch, out_dev_ptr, streams):
        """ceil(n/batch) asynchronous launches of `batch` boards, round-robin over the given cudaStream_t handles."""
        arr = (_vp * len(streams))(*streams)
        _ck(lib().cab_eval_dev_batches(self._h, codes_dev_ptr, n, batch, out_dev_ptr, arr, len(streams)))

    def predict_batch_bf16(self, codes):
        """bf16 tcgen05 forward (wide networks only: hidden widths multiples of 256)."""
        c = _as_codes(codes)
        out = np.zeros(len(c), dtype=np.float64)
        _ck(lib().cab_eval_bf16_host(self._h, c.reshape(-1), len(c), out))
        return out

    def prediction(self, codes, side_to_move, network_priors=False):
        """Decider::prediction (decider.cpp:31-61): (evaluation, [((from, to), logit), ...]) in the reference's map order,
        the position and the children that need the network evaluated in one batched launch."""
        c = _as_codes(codes)[0]
        ev, n = C.c_double(0), C.c_int(0)
        moves = (C.c_int * 512)()
        logits = np.zeros(256, dtype=np.float64)
        _ck(lib().cab_prediction_host(self._h, c, int(side_to_move), int(bool(network_priors)), C.byref(ev), moves, logits, 256, C.byref(n)))
        k = min(n.value, 256)
        return ev.value, [((moves[2 * i], moves[2 * i + 1]), float(logits[i])) for i in range(k)]

    def predict_batch_f16(self, codes):
        """Reduced-precision fused tcgen05 evaluation (fp16 operands) of a {64, N1, N2, N3, 1} net. Not a parity path."""
        c = _as_codes(codes)
        out = np.empty(len(c), dtype=np.float64)
        _ck(lib().cab_eval_f16_host(self._h, c.reshape(-1), len(c), out))
        return out

    def eval_f16_dev_ptr(self, codes_dev_ptr, n, out_dev_ptr, stream=0):
        _ck(lib().cab_eval_f16_dev(self._h, codes_dev_ptr, n, out_dev_ptr, stream))

    def eval_f16_dev_batches(self, codes_dev_ptr, n, batch, out_dev_ptr, streams):
        arr = (_vp * len(streams))(*streams)
        _ck(lib().cab_eval_f16_dev_batches(self._h, codes_dev_ptr, n, batch, out_dev_ptr, arr, len(streams)))

    def eval_bf16_dev_ptr(self, codes_dev_ptr, n, out_dev_ptr, stream=0):
        _ck(lib().cab_eval_bf16_dev(self._h, codes_dev_ptr, n, out_dev_ptr, stream))

    def forward_full(self, codes):
        """Network::propagate_for_training (network.cpp:139-153): activations of every layer, [B, sum(dims[1:])]."""
        c = _as_codes(codes)
        out = np.zeros((len(c), sum(self.dims[1:])), dtype=np.float64)
        _ck(lib().cab_forward_full_host(self._h, c.reshape(-1), len(c), out.reshape(-1)))
        return out

    def forward_thetas(self, codes):
        """propagate_for_training(unbounded) (network.cpp:139-153): (activations, pre-activations theta), each [B, sum(dims[1:])]."""
        c = _as_codes(codes)
        acts = np.zeros((len(c), sum(self.dims[1:])), dtype=np.float64)
        thetas = np.zeros_like(acts)
        _ck(lib().cab_forward_thetas_host(self._h, c.reshape(-1), len(c), acts.reshape(-1), thetas.reshape(-1)))
        return acts, thetas


# ---- network::Optimizer ---------------------------------------------------------------------------------------------
class Optimizer:
    """Mirror of network::Optimizer (network.h:142-156)."""

    DEFAULT_MAX_LEARNING_CONSTANT = 15.0
    DEFAULT_LEARNING_RATE = 1.1
    DEFAULT_INITIAL_LAMBDA = 2.0
    DEFAULT_TRAINING_REPETITIONS = 5

    @staticmethod
    def optimize(network, board, target, lam, max_const=15.0, rate=1.1):
        """One Optimizer::optimize step (network.cpp:223-279). Returns (new_lambda, accepted, E, E')."""
        c = board.codes() if isinstance(board, Board) else _as_codes(board)[0]
        c = np.ascontiguousarray(c, dtype=np.int8)
        lam_c, acc, e0, e1 = C.c_double(lam), C.c_int(0), C.c_double(0), C.c_double(0)
        _ck(lib().cab_train_case(network._h, c, target, C.byref(lam_c), max_const, rate, C.byref(acc), C.byref(e0), C.byref(e1)))
        return lam_c.value, bool(acc.value), e0.value, e1.value

    @staticmethod
    def dropout(network, rate=0.01, seed=0x5EED, offset=0):
        """Optimizer::dropout (network.cpp:281-292). Returns the number of re-randomised weights."""
        hits = C.c_long(0)
        _ck(lib().cab_dropout(network._h, rate, seed, offset, C.byref(hits)))
        return hits.value

    @staticmethod
    def train_sequence(network, codes, targets, lam=2.0, dropout_seed=0x5EED):
        """train_network's loop body over the cases in order (train.cpp:39-68). Returns (lambda, accepted[], resets)."""
        c = _as_codes(codes)
        t = np.ascontiguousarray(targets, dtype=np.float64)
        acc = np.zeros(len(c), dtype=np.uint8)
        lam_c, resets = C.c_double(lam), C.c_int(0)
        _ck(lib().cab_train_sequence(network._h, c.reshape(-1), t, len(c), C.byref(lam_c), dropout_seed, acc.ctypes.data, C.byref(resets)))
        return lam_c.value, acc, resets.value

    @staticmethod
    def train_batch(network, codes, targets, lam, max_const=15.0, rate=1.1):
        """Minibatch extension of optimize (mean update, one accept/reject). Returns (lambda, accepted, E, E')."""
        c = _as_codes(codes)
        t = np.ascontiguousarray(targets, dtype=np.float64)
        lam_c, acc, e0, e1 = C.c_double(lam), C.c_int(0), C.c_double(0), C.c_double(0)
        _ck(lib().cab_train_batch_host(network._h, c.reshape(-1), t, len(c), C.byref(lam_c), max_const, rate, C.byref(acc), C.byref(e0), C.byref(e1)))
        return lam_c.value, bool(acc.value), e0.value, e1.value

    @staticmethod
    def train_batch_bf16(network, codes, targets, lam, max_const=15.0, rate=1.1):
        """train_batch on the bf16 tcgen05 path (wide nets: hidden widths multiples of 256)."""
        c = _as_codes(codes)
        t = np.ascontiguousarray(targets, dtype=np.float64)
        lam_c, acc, e0, e1 = C.c_double(lam), C.c_int(0), C.c_double(0), C.c_double(0)
        _ck(lib().cab_train_batch_bf16_host(network._h, c.reshape(-1), t, len(c), C.byref(lam_c), max_const, rate, C.byref(acc), C.byref(e0), C.byref(e1)))
        return lam_c.value, bool(acc.value), e0.value, e1.value


class TrainStats(C.Structure):
    _fields_ = [("steps", C.c_long), ("accepted", C.c_long), ("resets", C.c_int), ("last_accepted", C.c_int), ("first_err", C.c_double),
                ("last_err", C.c_double), ("last_new_err", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def train_epoch_dev(network, codes_dev_ptr, targets_dev_ptr, n_cases, batch, lam, max_const=15.0, rate=1.1, reset_rule=True, dropout_seed=0x5EED,
                    trace=False, stream=0, bf16=False):
    """train_network's loop (train.cpp:39-68) over resident cases as consecutive minibatches, decisions on the device, one
    wait at the end. Returns (lambda, stats dict[, trace of (E, E') per step])."""
    st = TrainStats()
    steps = (n_cases + batch - 1) // batch
    tr = np.zeros((steps, 2)) if trace else None
    lam_c = C.c_double(lam)
    fn = lib().cab_train_epoch_bf16_dev if bf16 else lib().cab_train_epoch_dev
    _ck(fn(network._h, codes_dev_ptr, targets_dev_ptr, n_cases, batch, C.byref(lam_c), max_const, rate, 1 if reset_rule else 0, dropout_seed,
           C.byref(st), None if tr is None else tr.ctypes.data, stream))
    return (lam_c.value, st.as_dict(), tr) if trace else (lam_c.value, st.as_dict())


# ---- network::NetworkStorage ----------------------------------------------------------------------------------------
class NetworkStorage:
    """Mirror of network::NetworkStorage (network.h:158-183, network.cpp:294-381): process-wide current network,
    generation-rotating text checkpoints under network_dump/."""

    NETWORK_DUMP_DIRECTORY = "network_dump/"
    LATEST_NETWORK_FILE_PATH = NETWORK_DUMP_DIRECTORY + "latest.txt"
    SAVE_NETWORKS = False
    _current_network = None
    _network_count = 0
    _network_training_case = None

    @classmethod
    def current_network(cls):
        if cls._current_network is None:
            raise CabError(CAB_ESTATE, "NetworkStorage: no current network (network.cpp:300-303)")
        return cls._current_network

    @classmethod
    def initialize(cls, arg=None, device=0):
        """initialize() / initialize(dims) / initialize(path)  (network.cpp:311-343)."""
        if isinstance(arg, (str, bytes, os.PathLike)):
            cls._current_network = Network.loadFromFile(arg, device)
        else:
            cls._current_network = Network.randomNetwork(arg, device)
        cls._save_latest()
        return cls._current_network

    @classmethod
    def _save_latest(cls):
        if cls.SAVE_NETWORKS and cls._current_network is not None:
            os.makedirs(cls.NETWORK_DUMP_DIRECTORY, exist_ok=True)
            cls._current_network.dump(cls.LATEST_NETWORK_FILE_PATH)

    @classmethod
    def saveNetwork(cls):
        """network.cpp:345-355: latest.txt -> gen-<N>.txt, then write the new latest.txt."""
        if cls.SAVE_NETWORKS and cls._current_network is not None:
            if cls._network_count >= 0:
                past = cls.NETWORK_DUMP_DIRECTORY + "gen-%d.txt" % cls._network_count
                if os.path.exists(cls.LATEST_NETWORK_FILE_PATH):
                    os.replace(cls.LATEST_NETWORK_FILE_PATH, past)
            cls._network_count += 1
            cls._save_latest()

    @classmethod
    def flushStorage(cls):
        cls.saveNetwork()
        cls._current_network = None

    @classmethod
    def setTestCaseSelector(cls, selector):
        cls._network_training_case = selector

    @classmethod
    def saveBoard(cls, board, value):
        if cls._network_training_case is not None:
            cls._network_training_case(Board(board.records.copy()), value)


# ---- MCTS self-play resident on the GPU (csrc/cab_mcts.cu) ---------------------------------------------------------
class MctsParams(C.Structure):
    _fields_ = [("games", C.c_int), ("roots", C.c_int), ("playouts_per_move", C.c_int), ("in_flight", C.c_int), ("max_moves", C.c_int),
                ("network_priors", C.c_int), ("noise", C.c_int), ("precision", C.c_int), ("save_ratio", C.c_double), ("seed", C.c_uint64)]


class MctsStats(C.Structure):
    _fields_ = [(n, C.c_long) for n in ("playouts", "expansions", "boards_evaluated", "moves_played", "games_finished", "searches_done",
                                        "rounds", "arena_full", "eval_full", "cases_kept")] + [("seconds", C.c_double), ("evaluator_seconds", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class SelfPlay:
    """tree::MCTS self-play with virtual-loss leaf batching, trees and rules on the device (tree.cpp, decider.cpp:31-61)."""

    def __init__(self, net, games, playouts_per_move, roots=1, in_flight=8, max_moves=40, network_priors=0, noise=0, save_ratio=0.0, seed=1234,
                 precision=0):
        self.net = net
        self.params = MctsParams(games, roots, playouts_per_move, in_flight, max_moves, network_priors, noise, precision, save_ratio, seed)
        self._h = _vp(None)
        _ck(lib().cab_mcts_create(net._h, C.byref(self.params), C.byref(self._h)))

    def close(self):
        if self._h:
            lib().cab_mcts_destroy(self._h)
            self._h = _vp(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_positions(self, codes, side, no_capture=None):
        codes = np.ascontiguousarray(codes, dtype=np.int8).reshape(self.params.games, 64)
        side = np.ascontiguousarray(side, dtype=np.int8)
        nc = None if no_capture is None else np.ascontiguousarray(no_capture, dtype=np.int32)
        _ck(lib().cab_mcts_set_positions(self._h, codes, side, None if nc is None else nc.ctypes.data))

    def run(self, search_only=False, max_rounds=0):
        st = MctsStats()
        _ck(lib().cab_mcts_run(self._h, 1 if search_only else 0, max_rounds, C.byref(st)))
        return st.as_dict()

    def root_children(self, game, root=-1, cap=256):
        fr, to, vi = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        pr, vs = np.zeros(cap), np.zeros(cap)
        n, rv, rvs = C.c_int(0), C.c_int(0), C.c_double(0)
        _ck(lib().cab_mcts_root_children(self._h, game, root, cap, fr, to, pr, vi, vs, C.byref(n), C.byref(rv), C.byref(rvs)))
        k = min(n.value, cap)
        return {"from": fr[:k], "to": to[:k], "prior": pr[:k], "visits": vi[:k], "value_sum": vs[:k], "root_visits": rv.value,
                "root_value_sum": rvs.value}

    def game_state(self, game):
        codes = np.zeros(64, np.int8)
        side, over, result, moves = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        _ck(lib().cab_mcts_game_state(self._h, game, codes, C.byref(side), C.byref(over), C.byref(result), C.byref(moves)))
        return {"codes": codes, "side": side.value, "over": over.value, "result": result.value, "moves_played": moves.value}

    def cases(self):
        n = C.c_long(0)
        _ck(lib().cab_mcts_cases(self._h, None, None, 0, C.byref(n)))
        codes, targets = np.zeros((max(n.value, 1), 64), np.int8), np.zeros(max(n.value, 1))
        _ck(lib().cab_mcts_cases(self._h, codes.ctypes.data, targets.ctypes.data, n.value, C.byref(n)))
        return codes[:n.value], targets[:n.value]


def best_move(net, codes64, side, precision="fp64", tol=0.0):
    """First-ranked move under the network (argmax of net(child), first wins): (from, to, value, n_moves, n_refined)."""
    fr, to, nm, nr, val = C.c_int(-1), C.c_int(-1), C.c_int(0), C.c_int(0), C.c_double(0)
    _ck(lib().cab_best_move_host(net._h, np.ascontiguousarray(codes64, dtype=np.int8).reshape(64), int(side),
                                 {"fp64": 0, "f16": 1, "bf16": 2}[precision], float(tol), C.byref(fr), C.byref(to), C.byref(val), C.byref(nm),
                                 C.byref(nr)))
    return fr.value, to.value, val.value, nm.value, nr.value


def device_movegen(codes, side, queen_only=False, cap=256, no_reply=False, device=0):
    """Legal moves of n positions by the kernels' own rules (csrc/chess_rules.cuh), reference order."""
    codes = np.ascontiguousarray(codes, dtype=np.int8).reshape(-1, 64)
    side = np.ascontiguousarray(side, dtype=np.int8)
    n = len(codes)
    counts, moves = np.zeros(n, np.int32), np.zeros((n, cap, 3), np.uint8)
    flags = np.zeros((n, cap), np.uint8) if no_reply else None
    _ck(lib().cab_movegen_host(device, codes, side, n, 1 if queen_only else 0, cap, counts, moves.reshape(-1),
                               None if flags is None else flags.ctypes.data))
    return (counts, moves, flags) if no_reply else (counts, moves)
