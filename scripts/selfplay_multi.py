"""BASELINE.json configs[2] on N GPUs: one self-play process per GPU (independent games, replicated weights, no
collective; the search is resident on each GPU, csrc/cab_mcts.cu). Aggregate = total work / slowest process time.

    python scripts/selfplay_multi.py N_GPUS [games_per_gpu] [playouts] [moves] [in_flight_per_tree] [network_priors] [roots]
"""
import importlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(n, games=512, playouts=800, moves=20, in_flight=16, priors=1, roots=4):
    cab = importlib.import_module("chess-ai_b200")
    latest = os.path.join(tempfile.mkdtemp(), "latest.txt")
    cab.Network.from_weights([64, 144, 225, 100, 1], np.fromfile(os.path.join(ROOT, "tests", "golden", "latest_weights.f64"), dtype="<f8")).dump(latest)
    exe = os.path.join(ROOT, "chess-ai_b200", "lib", "selfplay")
    procs = [subprocess.Popen([exe, latest, str(games), str(playouts), str(moves), "-", str(in_flight), str(priors), str(d), str(1000 + d), "-", "0",
                               "fp64", str(roots)], stdout=subprocess.PIPE) for d in range(n)]
    outs = [json.loads(p.communicate()[0].decode().strip().splitlines()[-1]) for p in procs]
    assert all(p.returncode == 0 for p in procs)
    secs = max(o["seconds"] for o in outs)
    tot = {k: sum(o[k] for o in outs) for k in ("playouts", "boards_evaluated", "moves_played", "flushes", "gpu_launches")}
    return {"config": "configs[2]: MCTS self-play resident on the GPU, one process per GPU, no collective", "n_gpus": n, "host_cpus": os.cpu_count(),
            "games_per_gpu": games, "playouts_per_move": playouts, "moves_per_game": moves, "roots": roots, "in_flight_per_tree": in_flight,
            "network_priors": priors, "seconds_slowest": secs, "playouts_per_s": round(tot["playouts"] / secs, 1),
            "positions_evaluated_per_s": round(tot["boards_evaluated"] / secs, 1), "moves_played": tot["moves_played"],
            "avg_boards_per_flush": round(tot["boards_evaluated"] / max(1, tot["flushes"]), 1),
            "evaluator_busy_per_gpu": [round(o["evaluator_busy"], 3) for o in outs], "per_gpu_playouts_per_s": [round(o["playouts_per_s"], 1) for o in outs]}


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    print(json.dumps(run(*a)))
