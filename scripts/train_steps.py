"""A few minibatch training steps at B = 65536 (for ncu launch lists / timing of the training kernels)."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cab = importlib.import_module("chess-ai_b200")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
net = cab.Network.randomNetwork([64, 144, 225, 100, 1], seed=3)
net.set_weights(net.get_weights() * 0.1)
codes = cab.synthetic_boards(B)
targets = np.random.default_rng(0).uniform(-1, 1, B)
lam = 0.5
for i in range(6):
    t0 = time.perf_counter()
    lam, acc, e0, e1 = cab.Optimizer.train_batch(net, codes, targets, lam)
    print("step", i, "ms", round((time.perf_counter() - t0) * 1e3, 3), "accepted", acc, "E", e0, "->", e1)
