"""Throughput of the fused fp16 tcgen05 path (cab_eval_f16_dev) on the default net: launches of `batch` boards over 4
streams like bench.py's headline (inputs resident in HBM, > L2), plus one large launch. Prints JSON lines."""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cab = importlib.import_module("chess-ai_b200")
synth = importlib.import_module("chess-ai_b200.synth")

N = 1 << 22
w = np.fromfile(os.path.join(ROOT, "tests", "golden", "latest_weights.f64"), dtype="<f8")
net = cab.Network.from_weights([64, 144, 225, 100, 1], w)
codes = torch.from_numpy(np.tile(synth.synthetic_boards(1 << 20), (4, 1))).cuda()
out = torch.empty(N, dtype=torch.float64, device="cuda")
streams = [torch.cuda.Stream() for _ in range(4)]
for batch in [int(x) for x in sys.argv[1:]] or [4096, 16384, 65536, N]:
    def run():
        for i, b0 in enumerate(range(0, N, batch)):
            s = streams[i % 4]
            net.eval_f16_dev_ptr(codes.data_ptr() + b0 * 64, min(batch, N - b0), out.data_ptr() + b0 * 8, s.cuda_stream)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for _ in range(5):
        torch.cuda.synchronize()
        e0.record()
        for s in streams:
            s.wait_event(e0)
        run()
        for s in streams:
            e1.record(s) if s is streams[-1] else None
        torch.cuda.synchronize()
        t = torch.cuda.Event(enable_timing=True)
        times.append(None)
    # simple wall-clock timing around synchronize (launch overhead included, like a caller sees it)
    import time
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts))
    print(json.dumps({"config": "fused fp16 tcgen05 forward, default net 64-144-225-100-1 (latest.txt)", "boards": N, "batch": batch,
                      "streams": 4, "ms": round(dt * 1e3, 3), "evals_per_s": round(N / dt, 1),
                      "tflops_algorithmic": round(N * 128432 / dt / 1e12, 1)}), flush=True)
