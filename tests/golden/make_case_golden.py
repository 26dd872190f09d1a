"""Training-case golden vectors from the reference's OWN code (oracle/_ref: ref_case_target = initialization.cpp:221-227 +
make_cases.cpp:36-42; ref_save_case = Board::saveToFile + the target line of save_training_case) into
tests/golden/cases.npz. Run where /root/reference exists:
    python tests/golden/make_case_golden.py
"""
import ctypes as C
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libchessai_ref.so"))
i8p = np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")
L.ref_case_target.argtypes = [C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
L.ref_save_case.argtypes = [i8p, C.c_double, C.c_char_p]

rng = np.random.default_rng(20)
values = np.concatenate([[0.0, 1.0, 0.5, -0.25, 1.5, 0.1, 0.9999999999999999, 1e-300], rng.random(56), rng.normal(0.5, 0.6, 32)])
results = rng.integers(-1, 2, size=len(values)).astype(np.int32)
bounded, target = np.zeros(len(values)), np.zeros(len(values))
for i, (v, r) in enumerate(zip(values, results)):
    b, t = C.c_double(), C.c_double()
    assert L.ref_case_target(float(v), int(r), C.byref(b), C.byref(t)) == 0
    bounded[i], target[i] = b.value, t.value

games = np.load(os.path.join(HERE, "chess_games.npz"))
pick = rng.choice(len(games["codes"]), size=12, replace=False)
case_codes = np.ascontiguousarray(games["codes"][pick], dtype=np.int8)
case_targets = np.concatenate([target[:6], [-10.0, 10.0, 1.0 / 3.0, -7.2000000000000002, 5e-324, 123456.789]])
texts = []
with tempfile.TemporaryDirectory() as d:
    for i in range(len(pick)):
        path = os.path.join(d, "case_%d.txt" % i)
        assert L.ref_save_case(case_codes[i], float(case_targets[i]), path.encode()) == 0
        texts.append(open(path, "rb").read())
np.savez_compressed(os.path.join(HERE, "cases.npz"), values=values, results=results, bounded=bounded, target=target,
                    case_codes=case_codes, case_targets=case_targets,
                    case_texts=np.array(texts, dtype=object).astype("S"))
print("wrote cases.npz:", len(values), "target vectors,", len(texts), "case files; first text:\n", texts[0][:120].decode(), "...", texts[0][-40:])
