"""Decider::prediction golden vectors from the reference's OWN code (oracle/_ref: ref_prediction_codes, decider.cpp:31-61
with the shipped latest.txt) on positions of the recorded reference games, including every position from which a move
ends the game (the only case in which the shipped, inverted branch asks the network). Writes tests/golden/predictions.npz.
Run where /root/reference exists:   python tests/golden/make_prediction_golden.py
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libchessai_ref.so"))
i8p = np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
L.ref_net_load.restype = C.c_void_p
L.ref_net_load.argtypes = [C.c_char_p]
L.ref_prediction_codes.argtypes = [C.c_void_p, i8p, C.c_int, C.POINTER(C.c_double), i32p, i32p, f64p, C.c_int]
net = L.ref_net_load(b"/root/reference/network_dump/latest.txt")

g = np.load(os.path.join(HERE, "chess_games.npz"))
codes, side, game_of, overs, plies = g["codes"], g["side"], g["game_of"], g["overs"], g["plies"]
# position index of the last position BEFORE the final move of every finished game, plus a spread of ordinary ones
first = np.concatenate([[0], np.cumsum(plies + 1)[:-1]])
picks = [int(first[gi] + plies[gi] - 1) for gi in range(len(plies)) if overs[gi] and plies[gi] > 0]
rng = np.random.default_rng(5)
picks += [int(x) for x in rng.choice(len(codes), size=40, replace=False)]
# positions with a mating / stalemating move, found with the fast rules engine (selection only; the values come from _ref)
F = C.CDLL(os.path.join(ROOT, "chess-ai_b200", "lib", "libfastchess.so"))
F.fc_sizeof_position.restype = C.c_int
pos, child = (C.c_byte * F.fc_sizeof_position())(), (C.c_byte * F.fc_sizeof_position())()
buf, buf2 = (C.c_ubyte * (3 * 256))(), (C.c_ubyte * (3 * 256))()
nocap = g["nocap"]
found = []
for p in range(len(codes) - 1, -1, -1):
    if len(found) >= 24:
        break
    F.fc_set(pos, np.ascontiguousarray(codes[p]).ctypes.data_as(C.POINTER(C.c_byte)), int(side[p]), int(nocap[p]))
    k = F.fc_moves(pos, buf, 1)
    for i in range(k):
        C.memmove(child, pos, len(pos))
        F.fc_play(child, buf[3 * i], buf[3 * i + 1], buf[3 * i + 2])
        if F.fc_moves(child, buf2, 1) == 0:
            found.append(p)
            break
picks += found
picks = sorted(set(p for p in picks if p < len(codes)))

out = {"codes": [], "side": [], "eval": [], "n": [], "moves": [], "logits": []}
for p in picks:
    ev = C.c_double()
    fr, to, lg = np.zeros(256, np.int32), np.zeros(256, np.int32), np.zeros(256)
    k = L.ref_prediction_codes(net, np.ascontiguousarray(codes[p]), 1 if side[p] > 0 else 0, C.byref(ev), fr, to, lg, 256)
    assert 0 <= k <= 256
    out["codes"].append(codes[p]); out["side"].append(side[p]); out["eval"].append(ev.value); out["n"].append(k)
    out["moves"].append(np.stack([fr[:k], to[:k]], axis=1)); out["logits"].append(lg[:k])
n_net = sum(int((np.abs(np.abs(l) - 20.0) > 1e-12).sum()) for l in out["logits"])
np.savez_compressed(os.path.join(HERE, "predictions.npz"), codes=np.array(out["codes"], dtype=np.int8), side=np.array(out["side"], dtype=np.int8),
                    eval=np.array(out["eval"]), n=np.array(out["n"], dtype=np.int32), moves=np.concatenate(out["moves"]).astype(np.int32),
                    logits=np.concatenate(out["logits"]))
print("positions", len(picks), "moves", int(sum(out["n"])), "logits that came from the network", n_net)
