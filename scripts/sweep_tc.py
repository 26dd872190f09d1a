"""BASELINE.json configs[4]: widened MLP {64, 2048 x4, 1}, bf16 tcgen05 forward, batch sweep. Prints one JSON line per batch."""
import importlib, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cab = importlib.import_module("chess-ai_b200")
DIMS = [64, 2048, 2048, 2048, 2048, 1]
FLOP = 2 * (64 * 2048 + 3 * 2048 * 2048 + 2048)   # 25.43 MFLOP per board
net = cab.Network.randomNetwork(DIMS, seed=1)
net.set_weights(net.get_weights() * 0.05)
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"bf16_tflops": 1627.0, "bf16_tflops_sustained": 1384.0}
codes_all = torch.from_numpy(cab.synthetic_boards(65536)).cuda()
out = torch.empty(65536, dtype=torch.float64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for B in (256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536):
    for _ in range(3):
        net.eval_bf16_dev_ptr(codes_all.data_ptr(), B, out.data_ptr(), s)
    torch.cuda.synchronize()
    reps = max(3, min(50, (1 << 20) // B))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(3):
        e0.record()
        for _ in range(reps):
            net.eval_bf16_dev_ptr(codes_all.data_ptr(), B, out.data_ptr(), s)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    tf = FLOP * B / (best * 1e-3) * 1e-12
    print(json.dumps({"config": "widened MLP 64-2048x4-1 bf16 tcgen05 forward", "batch": B, "ms": round(best, 4), "evals_per_s": B / (best * 1e-3),
                      "tflops": round(tf, 1), "frac_of_bf16_burst_peak": round(tf / peaks["bf16_tflops"], 3),
                      "frac_of_bf16_sustained_peak": round(tf / peaks["bf16_tflops_sustained"], 3)}))
