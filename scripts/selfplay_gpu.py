#!/usr/bin/env python
"""configs[2] on one GPU: MCTS self-play resident on the device (cab_mcts_*). Prints one JSON line per configuration.
  python scripts/selfplay_gpu.py [games] [playouts] [moves] [roots] [in_flight] [network_priors] [device] [precision: 0 fp64 | 1 f16]"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cab = importlib.import_module("chess-ai_b200")
a = [int(x) for x in sys.argv[1:]]
games, playouts, moves, roots, in_flight, npri, device, precision = (a + [512, 800, 20, 1, 8, 1, 0, 0][len(a):])[:8]
w = np.fromfile(os.path.join(ROOT, "tests", "golden", "latest_weights.f64"), dtype="<f8")
net = cab.Network.from_weights([64, 144, 225, 100, 1], w, device=device)
sp = cab.SelfPlay(net, games=games, playouts_per_move=playouts, roots=roots, in_flight=in_flight, max_moves=moves, network_priors=npri, noise=1,
                  save_ratio=0.1, seed=1234 + device, precision=precision)
st = sp.run()
st.update({"games": games, "playouts_per_move": playouts, "max_moves": moves, "roots": roots, "in_flight": in_flight, "network_priors": npri, "precision": precision,
           "playouts_per_s": st["playouts"] / st["seconds"], "boards_per_s": st["boards_evaluated"] / st["seconds"],
           "evaluator_busy": st["evaluator_seconds"] / st["seconds"], "boards_per_round": st["boards_evaluated"] / max(st["rounds"], 1)})
print(json.dumps(st))
