"""Generates tests/golden/* by RUNNING THE REFERENCE'S OWN CODE (oracle/_ref/libchessai_ref.so, built by
oracle/Makefile from /root/reference/src in place) on the reference's shipped fixtures.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Everything written here is data (weights, boards, numbers); no reference source is copied.

Files
  latest_weights.f64   the 64 216 weights of network_dump/latest.txt, little-endian fp64, n,i,j order
  golden.npz           arrays (see keys below)
  golden.json          scalars / digests
  board_20.txt         one shipped training case verbatim (format fixture for the case parser)
  start_position.txt   assets/game_states/chess_default_start.txt verbatim (format fixture, no target line)
"""
import hashlib
import importlib
import json
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/"
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
from oracle_bind import PortOracle, RefOracle, build  # noqa: E402

cab = importlib.import_module("chess-ai_b200")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def fresh(code):
    """Run a snippet in a fresh interpreter so the reference's default-seeded mt19937 starts at its first draw."""
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "oracle"))
    out = subprocess.check_output([sys.executable, "-c", code], env=env)
    return json.loads(out.decode().strip().splitlines()[-1])


def main():
    build(ref=True)
    po, ro = PortOracle(), RefOracle()
    G, J = {}, {}
    latest = REF + "network_dump/latest.txt"
    rn = ro.load_net(latest)
    w = rn.get_weights()
    w.astype("<f8").tofile(os.path.join(HERE, "latest_weights.f64"))
    raw = open(latest, "rb").read()
    J["latest_txt"] = {"bytes": len(raw), "sha256": hashlib.sha256(raw).hexdigest(),
                       "fnv1a64": "%016x" % po.fnv1a64_file(latest), "dims": rn.dims, "n_weights": int(w.size),
                       "weights_sha256": sha(w), "max_abs": float(np.abs(w).max())}
    rn.dump("/tmp/_golden_ref_dump.txt")
    assert open("/tmp/_golden_ref_dump.txt", "rb").read() == raw, "reference dump is not byte-identical to latest.txt"

    # the 66 shipped training cases: piece inputs (Piece::code()), targets, forward outputs
    inputs, targets, codes = [], [], []
    for i in range(66):
        inp, t = ro.load_case(REF + "training_cases/board_%d.txt" % i)
        inputs.append(inp)
        targets.append(t)
        k = np.rint(inp * 2.5).astype(np.int8)
        assert np.array_equal(k.astype(np.float64) * 0.4, inp)
        codes.append(k)
    G["case_inputs"] = np.array(inputs)
    G["case_targets"] = np.array(targets)
    G["case_codes"] = np.array(codes, dtype=np.int8)
    G["case_forward"] = rn.predict(G["case_codes"])
    shutil.copy(REF + "training_cases/board_20.txt", os.path.join(HERE, "board_20.txt"))
    shutil.copy(REF + "assets/game_states/chess_default_start.txt", os.path.join(HERE, "start_position.txt"))

    # encoder table: all 19 code values through the reference's text parser + Piece::code()
    table = []
    for k in range(-9, 10):
        c = np.zeros(64, dtype=np.int8)
        c[5] = k
        table.append(ro.codes_to_inputs(c)[5])
    G["code_values"] = np.array(table)

    # Decider::prediction on the start position
    ev, logits = ro.load_net(latest).prediction_file(REF + "assets/game_states/chess_default_start.txt")
    J["start_prediction"] = {"eval": ev, "n_moves": int(len(logits)), "logits": sorted(set(float(x) for x in logits))}

    # synthetic boards (generator is ours; outputs are the reference's)
    syn = cab.synthetic_boards(2048)
    G["syn_codes"] = syn
    G["syn_forward_latest"] = rn.predict(syn)

    # one optimize step on board_20 (rejected) and the in-order epoch
    r1 = ro.load_net(latest)
    lam = r1.optimize(G["case_codes"][20], G["case_targets"][20], 2.0)
    J["optimize_board20"] = {"lambda_after": lam, "prediction_after": float(r1.predict(G["case_codes"][20:21])[0])}
    r2 = ro.load_net(latest)
    resets, lam = r2.train_sequence(G["case_codes"], G["case_targets"], 2.0)
    r2.dump("/tmp/_golden_epoch_dump.txt")
    J["epoch_latest"] = {"resets": resets, "final_lambda": lam, "weights_sha256": sha(r2.get_weights()),
                         "dump_fnv1a64": "%016x" % po.fnv1a64_file("/tmp/_golden_epoch_dump.txt"),
                         "dump_bytes": os.path.getsize("/tmp/_golden_epoch_dump.txt"),
                         "predict_board20": float(r2.predict(G["case_codes"][20:21])[0])}
    G["epoch_weights_sample"] = r2.get_weights()[::257].copy()

    # RNG stream + random networks + dropout from a FRESH process (default-seeded mt19937, math_util.h:35-36)
    J["rng"] = fresh(
        "import json,numpy as np\nfrom oracle_bind import RefOracle\nro=RefOracle()\n"
        "a=[ro.math_random() for _ in range(4)]\nprint(json.dumps({'first4':a}))")
    rnd = fresh(
        "import json,hashlib,numpy as np\nfrom oracle_bind import RefOracle\nro=RefOracle()\n"
        "n=ro.random_net([64,144,225,100,1]);w=n.get_weights()\n"
        "n2=ro.random_net([40,24]);w2=n2.get_weights()\n"
        "np.save('/tmp/_golden_rand_default.npy',w);np.save('/tmp/_golden_rand_small.npy',w2)\n"
        "print(json.dumps({'w0':w[0],'dims2':n2.dims}))")
    J["random_default"] = {"first_weight": rnd["w0"]}
    wr = np.load("/tmp/_golden_rand_default.npy")
    w_small = np.load("/tmp/_golden_rand_small.npy")
    small_dims = rnd["dims2"]
    J["random_small_dims"] = small_dims
    G["random_small_weights"] = w_small
    J["random_default"]["weights_sha256"] = sha(wr)
    G["random_default_weights_sample"] = wr[::101].copy()

    # forward of those random networks (reference outputs)
    rr = ro.net_from_weights([64, 144, 225, 100, 1], wr)
    G["syn_forward_random"] = rr.predict(syn[:512])
    rs = ro.net_from_weights(small_dims, w_small)
    G["syn_forward_small"] = rs.predict(syn[:512])

    # a training trace with accepts, rejects and lambda resets: random default net (stream continues into dropout),
    # 400 steps cycling over the 66 cases — all inside ONE fresh process so the mt19937 stream is reproducible.
    np.savez("/tmp/_golden_partial.npz", case_codes=G["case_codes"], case_targets=G["case_targets"])
    tr = fresh(
        "import json,hashlib,numpy as np\nfrom oracle_bind import RefOracle\nro=RefOracle()\n"
        "g=np.load('/tmp/_golden_partial.npz')\ncodes=g['case_codes'];t=g['case_targets']\n"
        "n=ro.random_net([64,144,225,100,1])\nlam=2.0;lams=[];res=0\n"
        "for s in range(400):\n"
        "    r,lam=n.train_sequence(codes[s%66:s%66+1],t[s%66:s%66+1],lam);res+=r;lams.append(lam)\n"
        "w=n.get_weights();np.save('/tmp/_golden_trace_w.npy',w)\n"
        "print(json.dumps({'lams':lams,'resets':res,'sha':hashlib.sha256(w.tobytes()).hexdigest()}))")
    G["trace_lambdas"] = np.array(tr["lams"])
    J["trace"] = {"steps": 400, "resets": tr["resets"], "weights_sha256": tr["sha"]}
    G["trace_weights_sample"] = np.load("/tmp/_golden_trace_w.npy")[::257].copy()

    np.savez_compressed(os.path.join(HERE, "golden.npz"), **G)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(J, f, indent=1, sort_keys=True)
    print(json.dumps(J, indent=1, sort_keys=True)[:3000])
    for fn in sorted(os.listdir(HERE)):
        print(fn, os.path.getsize(os.path.join(HERE, fn)))


if __name__ == "__main__":
    main()
