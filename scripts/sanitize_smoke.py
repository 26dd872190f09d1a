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
print("fused f16", net.predict_batch_f16(codes[:300])[:2])   # 3 tiles on 3 CTAs; then one CTA walking several tiles is not reachable at this size
import torch
dc = torch.from_numpy(codes).cuda()
dt = torch.from_numpy(t).cuda()
st = torch.cuda.Stream()
print("train_epoch_dev", cab.train_epoch_dev(student, dc.data_ptr(), dt.data_ptr(), 300, 128, 0.5, stream=st.cuda_stream))
sp = cab.SelfPlay(net, games=4, playouts_per_move=16, roots=2, in_flight=4, max_moves=2, network_priors=1, noise=1, save_ratio=0.5, seed=3)
print("selfplay", sp.run()["playouts"])
sp.close()
print("best_move", cab.best_move(net, np.array([4, 6, 7, 3, 1, 7, 6, 4] + [8] * 8 + [0] * 32 + [-8] * 8 + [-4, -6, -7, -3, -1, -7, -6, -4], dtype=np.int8), 1, "f16"))
print("ok")
