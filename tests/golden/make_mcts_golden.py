"""Golden vectors of the reference's OWN Monte-Carlo tree search for the GPU-resident search (csrc/cab_mcts.cu).

oracle/_ref: ref_mcts_root = Node + MCTS::expand_node + MCTS::mcts (tree.cpp:202-306) with the shipped latest.txt as the
Decider, no root noise (the noise comes from an un-vendored GSL, tree.cpp:318-347), SIMS playouts of 8 plies. Positions:
the start position, a spread of positions of the recorded reference games and positions from which a move ends the
game (the only case in which decider.cpp:52-57 as shipped asks the network for a child). Writes tests/golden/mcts_visits.npz:
per position the root children in std::map<Move> order (from, to, prior, visits, value_sum) and the root's own counts.
Run where /root/reference exists:   python tests/golden/make_mcts_golden.py
"""
import ctypes as C
import os
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SIMS = 150
L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libchessai_ref.so"))
i8p = np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
L.ref_net_load.restype = C.c_void_p
L.ref_net_load.argtypes = [C.c_char_p]
L.ref_mcts_root.argtypes = [C.c_void_p, i8p, C.c_int, C.c_int, C.c_int, C.c_int, i32p, i32p, f64p, i32p, f64p,
                            C.POINTER(C.c_int), C.POINTER(C.c_double)]
net = L.ref_net_load(b"/root/reference/network_dump/latest.txt")

g = np.load(os.path.join(HERE, "chess_games.npz"))
codes, side, nocap = g["codes"], g["side"], g["nocap"]
pred = np.load(os.path.join(HERE, "predictions.npz"))
# positions whose prediction has a network-driven logit (a move that ends the game)
off = np.concatenate([[0], np.cumsum(pred["n"])])
netty = [i for i in range(len(pred["n"])) if (np.abs(np.abs(pred["logits"][off[i]:off[i + 1]]) - 20.0) > 1e-12).any()]
rng = np.random.default_rng(17)
picks = [("game", int(i)) for i in sorted(rng.choice(len(codes), size=11, replace=False))]
# positions a few plies before the end of finished games: mates and stalemates inside the search horizon
first = np.concatenate([[0], np.cumsum(g["plies"] + 1)[:-1]])
late = [("game", int(first[gi] + g["plies"][gi] - back)) for gi in range(len(g["plies"])) if g["overs"][gi] and g["plies"][gi] > 6 for back in (2, 4)]
picks = [("game", 0)] + picks + [("pred", int(i)) for i in netty[:4]] + late[:16]

out = {k: [] for k in ("codes", "side", "nocap", "n", "from", "to", "prior", "visits", "value_sum", "root_visits", "root_value_sum")}
t0 = time.time()


def search(c, s, nc, q):
    fr, to = np.zeros(256, np.int32), np.zeros(256, np.int32)
    pr, vs = np.zeros(256), np.zeros(256)
    vi = np.zeros(256, np.int32)
    rv, rvs = C.c_int(), C.c_double()
    k = L.ref_mcts_root(net, c, 1 if s > 0 else 0, nc, SIMS, 256, fr, to, pr, vi, vs, C.byref(rv), C.byref(rvs))
    q.put((k, fr, to, pr, vi, vs, rv.value, rvs.value))


import multiprocessing as mp
ctx = mp.get_context("fork")
skipped = 0
for kind, i in picks:
    c = np.ascontiguousarray(codes[i] if kind == "game" else pred["codes"][i])
    s = int(side[i] if kind == "game" else pred["side"][i])
    nc = int(nocap[i]) if kind == "game" else 0
    # the reference's search DEBUG_ASSERTs (aborts) on some positions (a king-less board inside a playout): every
    # position is searched in a child process and the ones the reference itself cannot search are left out
    q = ctx.Queue()
    child = ctx.Process(target=search, args=(c, s, nc, q))
    with open(os.devnull, "w") as devnull:
        err = os.dup(2)
        os.dup2(devnull.fileno(), 2)
        child.start()
        os.dup2(err, 2)
        os.close(err)
    while child.is_alive() and q.empty():
        time.sleep(0.05)
    try:
        k, fr, to, pr, vi, vs, rv_value, rvs_value = q.get(timeout=2)
    except Exception:
        child.join()
        skipped += 1
        print(kind, i, "reference search aborted: skipped", flush=True)
        continue
    child.join()

    class _V:
        pass
    rv, rvs = _V(), _V()
    rv.value, rvs.value = rv_value, rvs_value
    assert 0 < k <= 256, (kind, i, k)
    out["codes"].append(c); out["side"].append(s); out["nocap"].append(nc); out["n"].append(k)
    out["from"].append(fr[:k]); out["to"].append(to[:k]); out["prior"].append(pr[:k]); out["visits"].append(vi[:k]); out["value_sum"].append(vs[:k])
    out["root_visits"].append(rv.value); out["root_value_sum"].append(rvs.value)
    print(kind, i, "children", k, "root visits", rv.value, "max child visits", int(vi[:k].max()), "%.1fs" % (time.time() - t0), flush=True)
np.savez_compressed(os.path.join(HERE, "mcts_visits.npz"), sims=np.int32(SIMS), codes=np.array(out["codes"], dtype=np.int8),
                    side=np.array(out["side"], dtype=np.int8), nocap=np.array(out["nocap"], dtype=np.int32), n=np.array(out["n"], dtype=np.int32),
                    frm=np.concatenate(out["from"]).astype(np.int32), to=np.concatenate(out["to"]).astype(np.int32),
                    prior=np.concatenate(out["prior"]), visits=np.concatenate(out["visits"]).astype(np.int32),
                    value_sum=np.concatenate(out["value_sum"]), root_visits=np.array(out["root_visits"], dtype=np.int32),
                    root_value_sum=np.array(out["root_value_sum"]))
print("wrote mcts_visits.npz:", len(out["n"]), "positions;", skipped, "skipped (reference aborts)")
