"""ctypes bindings for the two CPU checkers (TEST INFRASTRUCTURE — see oracle/mlp_oracle.c header).

  PortOracle : oracle/libmlp_oracle.so       — the plain-C restatement (always buildable: gcc only)
  RefOracle  : oracle/_ref/libchessai_ref.so — the reference's own translation units (prebuilt where
               /root/reference exists; travels to the GPU box as a built file)

Only tests/, bench.py's cpu_baseline / --impl reference leg and __graft_entry__.smoke() import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libmlp_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libchessai_ref.so")
REF_SO_O0 = os.path.join(HERE, "_ref", "libchessai_ref_O0.so")   # the same translation units at the reference's as-shipped flags (no -O)

_i8p = np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(ref=True):
    """Compile the checkers (the C port always; the reference library only when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "libmlp_oracle.so"])
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _codes(codes):
    a = np.ascontiguousarray(codes, dtype=np.int8)
    assert a.shape[-1] == 64
    return a.reshape(-1, 64)


class PortNet:
    """OrcNet handle of oracle/mlp_oracle.c."""

    def __init__(self, lib, ptr):
        self.lib, self.ptr = lib, ptr

    def __del__(self):
        if getattr(self, "ptr", None):
            self.lib.orc_net_free(self.ptr)
            self.ptr = None

    @property
    def dims(self):
        out = np.zeros(self.lib.orc_net_num_layers(self.ptr), dtype=np.int32)
        self.lib.orc_net_dims(self.ptr, out)
        return [int(x) for x in out]

    @property
    def num_weights(self):
        return int(self.lib.orc_net_num_weights(self.ptr))

    def get_weights(self):
        out = np.zeros(self.num_weights, dtype=np.float64)
        self.lib.orc_net_get_weights(self.ptr, out)
        return out

    def set_weights(self, w):
        w = np.ascontiguousarray(w, dtype=np.float64)
        assert w.size == self.num_weights
        self.lib.orc_net_set_weights(self.ptr, w)

    def clone(self):
        return PortNet(self.lib, self.lib.orc_net_clone(self.ptr))

    def predict(self, codes):
        c = _codes(codes)
        out = np.zeros(len(c), dtype=np.float64)
        self.lib.orc_predict(self.ptr, c, len(c), out)
        return out

    def time_predict(self, codes):
        c = _codes(codes)
        out = np.zeros(len(c), dtype=np.float64)
        return float(self.lib.orc_time_predict(self.ptr, c, len(c), out)), out

    def forward_full(self, code64):
        c = _codes(code64)
        n = sum(self.dims)
        ne, th = np.zeros(n), np.zeros(n)
        self.lib.orc_forward_full(self.ptr, c, ne, th)
        return ne, th

    def optimize(self, code64, target, lam, max_const=15.0, rate=1.1):
        c = _codes(code64)
        lam_c, e0, e1 = C.c_double(lam), C.c_double(0), C.c_double(0)
        acc = self.lib.orc_optimize(self.ptr, c, target, C.byref(lam_c), max_const, rate, C.byref(e0), C.byref(e1))
        return int(acc), lam_c.value, e0.value, e1.value

    def train_batch(self, codes, targets, lam, max_const=15.0, rate=1.1):
        c = _codes(codes)
        t = np.ascontiguousarray(targets, dtype=np.float64)
        lam_c, e0, e1 = C.c_double(lam), C.c_double(0), C.c_double(0)
        acc = self.lib.orc_train_batch(self.ptr, c, t, len(c), C.byref(lam_c), max_const, rate, C.byref(e0), C.byref(e1))
        return int(acc), lam_c.value, e0.value, e1.value

    def train_sequence(self, codes, targets, lam, mt=None, philox_seed=None):
        c = _codes(codes)
        t = np.ascontiguousarray(targets, dtype=np.float64)
        lam_c = C.c_double(lam)
        accepted = np.zeros(len(c), dtype=np.uint8)
        resets = self.lib.orc_train_sequence(self.ptr, c, t, len(c), C.byref(lam_c), mt, 0 if philox_seed is None else 1,
                                             C.c_uint64(philox_seed or 0), accepted)
        return int(resets), lam_c.value, accepted

    def time_train(self, codes, targets, lam):
        c = _codes(codes)
        t = np.ascontiguousarray(targets, dtype=np.float64)
        lam_c = C.c_double(lam)
        return float(self.lib.orc_time_train(self.ptr, c, t, len(c), C.byref(lam_c))), lam_c.value

    def dropout_mt(self, rate, mt):
        return int(self.lib.orc_dropout_mt(self.ptr, rate, mt))

    def dropout_philox(self, rate, seed, offset):
        return int(self.lib.orc_dropout_philox(self.ptr, rate, C.c_uint64(seed), C.c_uint64(offset)))

    def dump(self, path):
        assert self.lib.orc_net_dump(self.ptr, path.encode()) == 0


class PortOracle:
    def __init__(self, path=PORT_SO):
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = C.CDLL(path)
        vp = C.c_void_p
        L.orc_piece_input.restype = C.c_double
        L.orc_piece_input.argtypes = [C.c_int] * 3
        L.orc_piece_code.restype = C.c_int
        L.orc_piece_code.argtypes = [C.c_int] * 3
        L.orc_code_input.restype = C.c_double
        L.orc_code_input.argtypes = [C.c_int]
        L.orc_encode_records.argtypes = [_u8p, C.c_long, _i8p]
        L.orc_f.restype = C.c_double
        L.orc_f.argtypes = [C.c_double]
        L.orc_df.restype = C.c_double
        L.orc_df.argtypes = [C.c_double]
        L.orc_normalize_dims.restype = C.c_int
        L.orc_normalize_dims.argtypes = [_i32p, C.c_int, _i32p]
        L.orc_net_new.restype = vp
        L.orc_net_new.argtypes = [_i32p, C.c_int]
        L.orc_net_free.argtypes = [vp]
        L.orc_net_clone.restype = vp
        L.orc_net_clone.argtypes = [vp]
        L.orc_net_num_weights.restype = C.c_long
        L.orc_net_num_weights.argtypes = [vp]
        L.orc_net_num_layers.restype = C.c_int
        L.orc_net_num_layers.argtypes = [vp]
        L.orc_net_dims.argtypes = [vp, _i32p]
        L.orc_net_get_weights.argtypes = [vp, _f64p]
        L.orc_net_set_weights.argtypes = [vp, _f64p]
        L.orc_predict.argtypes = [vp, _i8p, C.c_long, _f64p]
        L.orc_time_predict.restype = C.c_double
        L.orc_time_predict.argtypes = [vp, _i8p, C.c_long, _f64p]
        L.orc_forward_full.argtypes = [vp, _i8p, _f64p, _f64p]
        L.orc_optimize.restype = C.c_int
        L.orc_optimize.argtypes = [vp, _i8p, C.c_double, C.POINTER(C.c_double), C.c_double, C.c_double,
                                   C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_train_batch.restype = C.c_int
        L.orc_train_batch.argtypes = [vp, _i8p, _f64p, C.c_long, C.POINTER(C.c_double), C.c_double, C.c_double,
                                      C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_train_sequence.restype = C.c_int
        L.orc_train_sequence.argtypes = [vp, _i8p, _f64p, C.c_long, C.POINTER(C.c_double), vp, C.c_int, C.c_uint64, _u8p]
        L.orc_time_train.restype = C.c_double
        L.orc_time_train.argtypes = [vp, _i8p, _f64p, C.c_long, C.POINTER(C.c_double)]
        L.orc_mt_new.restype = vp
        L.orc_mt_new.argtypes = [C.c_uint32]
        L.orc_mt_free.argtypes = [vp]
        L.orc_mt_next.restype = C.c_uint32
        L.orc_mt_next.argtypes = [vp]
        L.orc_mt_canonical.restype = C.c_double
        L.orc_mt_canonical.argtypes = [vp]
        L.orc_random.restype = C.c_double
        L.orc_random.argtypes = [vp, C.c_double, C.c_double]
        L.orc_random_network.argtypes = [vp, vp]
        L.orc_dropout_mt.restype = C.c_long
        L.orc_dropout_mt.argtypes = [vp, C.c_double, vp]
        L.orc_dropout_philox.restype = C.c_long
        L.orc_dropout_philox.argtypes = [vp, C.c_double, C.c_uint64, C.c_uint64]
        L.orc_philox4x32_10.argtypes = [C.c_uint32] * 6 + [np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")]
        L.orc_net_dump.restype = C.c_int
        L.orc_net_dump.argtypes = [vp, C.c_char_p]
        L.orc_net_load.restype = vp
        L.orc_net_load.argtypes = [C.c_char_p]
        L.orc_fnv1a64_file.restype = C.c_uint64
        L.orc_fnv1a64_file.argtypes = [C.c_char_p]
        L.orc_prior_softmax.argtypes = [_f64p, C.c_int, C.c_double, _f64p]
        L.orc_puct.restype = C.c_double
        L.orc_puct.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
        L.orc_parse_case_file.restype = C.c_int
        L.orc_parse_case_file.argtypes = [C.c_char_p, _u8p, C.POINTER(C.c_double), C.POINTER(C.c_int)]

    # -- nets
    def new_net(self, dims):
        d = np.ascontiguousarray(dims, dtype=np.int32)
        return PortNet(self.lib, self.lib.orc_net_new(d, len(d)))

    def load_net(self, path):
        p = self.lib.orc_net_load(path.encode())
        if not p:
            raise IOError("cannot load " + path)
        return PortNet(self.lib, p)

    def net_from_weights(self, dims, w):
        n = self.new_net(dims)
        n.set_weights(w)
        return n

    def normalize_dims(self, dims):
        d = np.ascontiguousarray(dims, dtype=np.int32)
        out = np.zeros(len(d) + 5, dtype=np.int32)
        m = self.lib.orc_normalize_dims(d, len(d), out)
        return [int(x) for x in out[:m]]

    def random_net(self, dims, mt):
        n = self.new_net(dims)
        self.lib.orc_random_network(n.ptr, mt)
        return n

    # -- rng
    def mt_new(self, seed=5489):
        return self.lib.orc_mt_new(seed)

    def philox(self, ctr, key):
        out = np.zeros(4, dtype=np.uint32)
        self.lib.orc_philox4x32_10(*[int(x) for x in ctr], *[int(x) for x in key], out)
        return out

    # -- search math
    def prior_softmax(self, logits, color_multiplier):
        l = np.ascontiguousarray(logits, dtype=np.float64)
        out = np.zeros_like(l)
        self.lib.orc_prior_softmax(l, len(l), color_multiplier, out)
        return out

    # -- encoder / formats
    def encode_records(self, records):
        r = np.ascontiguousarray(records, dtype=np.uint8).reshape(-1, 64, 3)
        out = np.zeros((len(r), 64), dtype=np.int8)
        self.lib.orc_encode_records(r, len(r), out)
        return out

    def parse_case_file(self, path):
        rec = np.zeros((64, 3), dtype=np.uint8)
        t, has = C.c_double(0), C.c_int(0)
        rc = self.lib.orc_parse_case_file(path.encode(), rec, C.byref(t), C.byref(has))
        if rc != 0:
            raise IOError("bad case file %s (rc=%d)" % (path, rc))
        return rec, (t.value if has.value else None)

    def fnv1a64_file(self, path):
        return int(self.lib.orc_fnv1a64_file(path.encode()))


class RefNet:
    def __init__(self, lib, ptr):
        self.lib, self.ptr = lib, ptr

    def __del__(self):
        if getattr(self, "ptr", None):
            self.lib.ref_net_free(self.ptr)
            self.ptr = None

    @property
    def dims(self):
        out = np.zeros(self.lib.ref_net_num_layers(self.ptr), dtype=np.int32)
        self.lib.ref_net_dims(self.ptr, out)
        return [int(x) for x in out]

    @property
    def num_weights(self):
        return int(self.lib.ref_net_num_weights(self.ptr))

    def get_weights(self):
        out = np.zeros(self.num_weights, dtype=np.float64)
        self.lib.ref_net_get_weights(self.ptr, out)
        return out

    def set_weights(self, w):
        w = np.ascontiguousarray(w, dtype=np.float64)
        assert w.size == self.num_weights
        self.lib.ref_net_set_weights(self.ptr, w)

    def clone(self):
        return RefNet(self.lib, self.lib.ref_net_clone(self.ptr))

    def predict(self, codes):
        c = _codes(codes)
        out = np.zeros(len(c), dtype=np.float64)
        assert self.lib.ref_predict_codes(self.ptr, c, len(c), out) == 0
        return out

    def time_predict(self, codes, threads=1, reps=1):
        c = _codes(codes)
        out = np.zeros(len(c), dtype=np.float64)
        return float(self.lib.ref_time_predict(self.ptr, c, len(c), threads, reps, out)), out

    def optimize(self, code64, target, lam, max_const=15.0, rate=1.1):
        c = _codes(code64)
        lam_c = C.c_double(lam)
        assert self.lib.ref_optimize_codes(self.ptr, c, target, C.byref(lam_c), max_const, rate) == 0
        return lam_c.value

    def train_sequence(self, codes, targets, lam):
        c = _codes(codes)
        t = np.ascontiguousarray(targets, dtype=np.float64)
        lam_c = C.c_double(lam)
        resets = self.lib.ref_train_sequence(self.ptr, c, t, len(c), C.byref(lam_c))
        return int(resets), lam_c.value

    def time_train(self, codes, targets, lam):
        c = _codes(codes)
        t = np.ascontiguousarray(targets, dtype=np.float64)
        lam_c = C.c_double(lam)
        return float(self.lib.ref_time_train(self.ptr, c, t, len(c), C.byref(lam_c))), lam_c.value

    def dropout(self, rate=0.01):
        self.lib.ref_dropout(self.ptr, rate)

    def dump(self, path):
        assert self.lib.ref_net_dump(self.ptr, path.encode()) == 0

    def prediction_file(self, board_path, white_to_move=True, cap=256):
        ev = C.c_double(0)
        logits = np.zeros(cap, dtype=np.float64)
        k = self.lib.ref_prediction_file(self.ptr, board_path.encode(), 1 if white_to_move else 0, C.byref(ev), logits, cap)
        return ev.value, logits[:min(k, cap)].copy()


class RefOracle:
    """The reference's own code. Raises FileNotFoundError when oracle/_ref was never built."""

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self, path=REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = self.lib = C.CDLL(path)
        vp = C.c_void_p
        L.ref_net_load.restype = vp
        L.ref_net_load.argtypes = [C.c_char_p]
        L.ref_net_new.restype = vp
        L.ref_net_new.argtypes = [_i32p, C.c_int]
        L.ref_net_random.restype = vp
        L.ref_net_random.argtypes = [_i32p, C.c_int]
        L.ref_net_clone.restype = vp
        L.ref_net_clone.argtypes = [vp]
        L.ref_net_free.argtypes = [vp]
        L.ref_net_num_layers.restype = C.c_int
        L.ref_net_num_layers.argtypes = [vp]
        L.ref_net_dims.argtypes = [vp, _i32p]
        L.ref_net_num_weights.restype = C.c_long
        L.ref_net_num_weights.argtypes = [vp]
        L.ref_net_get_weights.argtypes = [vp, _f64p]
        L.ref_net_set_weights.argtypes = [vp, _f64p]
        L.ref_net_dump.restype = C.c_int
        L.ref_net_dump.argtypes = [vp, C.c_char_p]
        L.ref_load_case.restype = C.c_int
        L.ref_load_case.argtypes = [C.c_char_p, _f64p, C.POINTER(C.c_double)]
        L.ref_codes_to_inputs.restype = C.c_int
        L.ref_codes_to_inputs.argtypes = [_i8p, _f64p]
        L.ref_predict_codes.restype = C.c_int
        L.ref_predict_codes.argtypes = [vp, _i8p, C.c_long, _f64p]
        L.ref_time_predict.restype = C.c_double
        L.ref_time_predict.argtypes = [vp, _i8p, C.c_long, C.c_int, C.c_int, _f64p]
        L.ref_optimize_codes.restype = C.c_int
        L.ref_optimize_codes.argtypes = [vp, _i8p, C.c_double, C.POINTER(C.c_double), C.c_double, C.c_double]
        L.ref_train_sequence.restype = C.c_int
        L.ref_train_sequence.argtypes = [vp, _i8p, _f64p, C.c_long, C.POINTER(C.c_double)]
        L.ref_time_train.restype = C.c_double
        L.ref_time_train.argtypes = [vp, _i8p, _f64p, C.c_long, C.POINTER(C.c_double)]
        L.ref_dropout.argtypes = [vp, C.c_double]
        L.ref_math_random.restype = C.c_double
        L.ref_prediction_file.restype = C.c_int
        L.ref_prediction_file.argtypes = [vp, C.c_char_p, C.c_int, C.POINTER(C.c_double), _f64p, C.c_int]

    def load_net(self, path):
        p = self.lib.ref_net_load(path.encode())
        if not p:
            raise IOError("cannot load " + path)
        return RefNet(self.lib, p)

    def new_net(self, dims):
        d = np.ascontiguousarray(dims, dtype=np.int32)
        return RefNet(self.lib, self.lib.ref_net_new(d, len(d)))

    def random_net(self, dims):
        d = np.ascontiguousarray(dims, dtype=np.int32)
        return RefNet(self.lib, self.lib.ref_net_random(d, len(d)))

    def net_from_weights(self, dims, w):
        n = self.new_net(dims)
        n.set_weights(w)
        return n

    def load_case(self, path):
        inputs = np.zeros(64, dtype=np.float64)
        t = C.c_double(0)
        self.lib.ref_load_case(path.encode(), inputs, C.byref(t))
        return inputs, t.value

    def codes_to_inputs(self, code64):
        c = _codes(code64)
        out = np.zeros(64, dtype=np.float64)
        assert self.lib.ref_codes_to_inputs(c, out) == 0
        return out

    def math_random(self):
        return float(self.lib.ref_math_random())
