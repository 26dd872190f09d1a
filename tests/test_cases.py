"""Training-case pipeline (chess-ai_b200/host/cases.hpp): target rule, case text format and packed
shards against vectors produced by the reference's own code (tests/golden/make_case_golden.py -> cases.npz)."""
import ctypes as C
import importlib
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "chess-ai_b200", "lib", "libfastchess.so")
i8p = np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def fc():
    if not os.path.exists(LIB):
        importlib.import_module("chess-ai_b200.build").build()
    L = C.CDLL(LIB)
    L.fc_case_target.argtypes = [C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.fc_case_text.argtypes = [i8p, C.c_double, C.c_char_p, C.c_int]
    L.fc_case_parse.argtypes = [C.c_char_p, i8p, C.POINTER(C.c_double)]
    L.fc_shard_append.argtypes = [C.c_char_p, i8p, f64p, C.c_long]
    L.fc_shard_read.argtypes = [C.c_char_p, i8p, f64p, C.c_long]
    L.fc_shard_read.restype = C.c_long
    return L


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(ROOT, "tests", "golden", "cases.npz"))


def test_target_rule_bit_exact(fc, g):
    for v, r, b_ref, t_ref in zip(g["values"], g["results"], g["bounded"], g["target"]):
        b, t = C.c_double(), C.c_double()
        fc.fc_case_target(float(v), int(r), C.byref(b), C.byref(t))
        assert b.value == b_ref and t.value == t_ref, (v, r)
    assert g["target"].min() >= -10.0 and g["target"].max() <= 10.0


def test_case_text_is_byte_identical_to_the_reference_writer(fc, g):
    buf = C.create_string_buffer(4096)
    for codes, target, ref_text in zip(g["case_codes"], g["case_targets"], g["case_texts"]):
        n = fc.fc_case_text(np.ascontiguousarray(codes), float(target), buf, 4096)
        assert n > 0 and buf.raw[:n] == bytes(ref_text)
        back, t = np.zeros(64, dtype=np.int8), C.c_double()
        assert fc.fc_case_parse(bytes(ref_text), back, C.byref(t)) == 0
        assert (back == codes).all() and t.value == target           # %.17e round-trips every double


def test_case_text_matches_the_python_mirror(fc, g):
    cab = importlib.import_module("chess-ai_b200")
    b, target = cab.Board.from_text(bytes(g["case_texts"][0]).decode())
    assert (np.asarray(cab.Board.from_codes(g["case_codes"][0]).records) == np.asarray(b.records)).all()
    assert target == g["case_targets"][0]


def test_shipped_training_case_parses(fc):
    text = open(os.path.join(ROOT, "tests", "golden", "board_20.txt"), "rb").read()
    codes, t = np.zeros(64, dtype=np.int8), C.c_double()
    assert fc.fc_case_parse(text, codes, C.byref(t)) == 0
    assert t.value == 0.5245862881359199 and np.count_nonzero(codes) == 32 and sorted(np.abs(codes[codes != 0]))[:2] == [1, 1]
    assert fc.fc_case_parse(b"8 8\n0 - wizard white .\n", codes, C.byref(t)) == 1       # malformed -> error, not garbage


def test_packed_shards_round_trip_and_concatenate(fc, g, tmp_path):
    path = str(tmp_path / "cases.shard").encode()
    codes, targets = np.ascontiguousarray(g["case_codes"]), np.ascontiguousarray(g["case_targets"], dtype=np.float64)
    assert fc.fc_shard_append(path, codes[:5], targets[:5], 5) == 0
    assert fc.fc_shard_append(path, codes[5:], targets[5:], len(codes) - 5) == 0
    assert os.path.getsize(path) == 72 * len(codes)
    rec = np.fromfile(path.decode(), dtype=np.dtype([("codes", "i1", 64), ("target", "<f8")]))   # the documented layout
    assert (rec["codes"] == codes).all() and (rec["target"].view(np.uint64) == targets.view(np.uint64)).all()
    c2, t2 = np.zeros_like(codes), np.zeros_like(targets)
    assert fc.fc_shard_read(path, c2, t2, len(codes)) == len(codes)
    assert (c2 == codes).all() and (t2.view(np.uint64) == targets.view(np.uint64)).all()
    with open(path.decode(), "ab") as f:
        f.write(b"x")                                                                             # torn tail
    assert fc.fc_shard_read(path, c2, t2, len(codes)) == -1
    assert fc.fc_shard_read(b"/nonexistent/shard", c2, t2, 1) == -1
