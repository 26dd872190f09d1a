"""Small end-to-end pass over every kernel family, sized for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cab = importlib.import_module("chess-ai_b200")
synth = importlib.import_module("chess-ai_b200.synth")

w = np.fromfile(os.path.join(ROOT, "tests", "golden", "latest_weights.f64"), dtype="<f8")
dims = [64, 144, 225, 100, 1]
net = cab.Network.from_weights(dims, w)
codes = synth.synthetic_boards(300)
y = net.predict_batch(codes)                                  # forward, ragged batch
print("forward", y[:2])
full = net.forward_full(codes[:9])                            # forward with saved activations
student = cab.Network.randomNetwork(dims, seed=5)
student.set_weights(student.get_weights() * 0.1)
t = np.clip(y, -3.9, 3.9)
print("train_batch", cab.Optimizer.train_batch(student, codes, t, 0.5))       # forward+save, backward, grad, update, re-forward
print("train_sequence", cab.Optimizer.train_sequence(student, codes[:5], t[:5], lam=2.0)[0])
print("optimize", cab.Optimizer.optimize(student, codes[0], float(t[0]), 2.0))
print("dropout", cab.Optimizer.dropout(student, 0.01, seed=1, offset=0))
wide = [64, 256, 256, 1]
rng = np.random.default_rng(0)
ww = np.concatenate([(rng.uniform(-1, 1, (k, n)) * 2 / np.sqrt(k)).reshape(-1) for k, n in zip(wide[:-1], wide[1:])])
wnet = cab.Network.from_weights(wide, ww)
print("bf16 forward", wnet.predict_batch_bf16(codes[:130])[:2])
print("bf16 train", cab.Optimizer.train_batch_bf16(wnet, codes[:200], t[:200], 2e-3))
rec = synth.codes_to_records(codes[:7])
print("encode", np.array_equal(cab.encode_records(rec), codes[:7]))
print("ok")
