"""Worker of tests/test_tc_gpu.py::test_cta_pair_kernel_matches_single_cta: runs the bf16 forward and one bf16 training
step with the GEMM kind forced by CAB_TC_PAIR (read once per process) and saves the results."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cab = importlib.import_module("chess-ai_b200")
synth = importlib.import_module("chess-ai_b200.synth")

dims = [64, 256, 512, 256, 1]
rng = np.random.default_rng(42)
w = np.concatenate([(rng.uniform(-1, 1, (k, n)) * 2.0 / np.sqrt(k)).reshape(-1) for k, n in zip(dims[:-1], dims[1:])])
codes = synth.synthetic_boards(777, seed=9)          # ragged: 3 full 256-row tiles + 9 rows
targets = np.clip(rng.normal(0, 1.5, len(codes)), -3.9, 3.9)
net = cab.Network.from_weights(dims, w)
y = net.predict_batch_bf16(codes)
lam, acc, e0, e1 = cab.Optimizer.train_batch_bf16(net, codes, targets, 2e-3)
np.savez(sys.argv[1], y=y, w_after=net.get_weights(), w_before=w, stats=np.array([lam, float(acc), e0, e1]))
