"""BASELINE.json configs[4]: widened MLP {64,2048,2048,2048,2048,1}, bf16 tcgen05 forward+backward (cab_grad_bf16_dev),
batch sweep. Prints one JSON line per batch: TFLOP/s over the algorithmic 76.0 MFLOP/board (fwd 2P + dX + dW, layer-0
dX skipped) and the share of the measured bf16 peaks. Inputs resident in HBM; CUDA events; L2 flushed
between timed iterations."""
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

DIMS = [64, 2048, 2048, 2048, 2048, 1]
P_GEMM = sum(k * n for k, n in zip(DIMS[:-1], DIMS[1:]))
FWD = 2 * P_GEMM
BWD_DX = 2 * sum(k * n for k, n in list(zip(DIMS[:-1], DIMS[1:]))[1:])
BWD_DW = 2 * P_GEMM
FLOP_PER_BOARD = FWD + BWD_DX + BWD_DW
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}


def main():
    batches = [int(x) for x in sys.argv[1:]] or [256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536]
    rng = np.random.default_rng(1)
    w = np.concatenate([(rng.uniform(-1, 1, (k, n)) * 2.0 / np.sqrt(k)).reshape(-1) for k, n in zip(DIMS[:-1], DIMS[1:])])
    net = cab.Network.from_weights(DIMS, w)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()
    for B in batches:
        codes = torch.from_numpy(synth.synthetic_boards(B, seed=3)).cuda()
        targets = torch.randn(B, dtype=torch.float64, device="cuda").clamp_(-3.9, 3.9)
        L = cab.lib()
        def step():
            rc = L.cab_grad_bf16_dev(net._h, codes.data_ptr(), targets.data_ptr(), B, stream.cuda_stream)
            assert rc == 0, cab.lib().cab_last_error()
        with torch.cuda.stream(stream):
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            times = []
            for _ in range(8):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                step()
                e1.record(stream)
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        tf = FLOP_PER_BOARD * B / (ms * 1e-3) / 1e12
        print(json.dumps({"config": "widened MLP bf16 tcgen05 fwd+bwd", "dims": DIMS, "batch": B, "ms": round(ms, 4),
                          "boards_per_s": round(B / (ms * 1e-3), 1), "tflops": round(tf, 1),
                          "flop_per_board": FLOP_PER_BOARD, "includes": "grad memset (102 MB) + codes->bf16 + 4 fwd GEMM + gemv + out/psi kernels + 4 dW GEMM + 3 dX GEMM",
                          "frac_bf16_burst": round(tf / 1627.0, 3), "frac_bf16_sustained": round(tf / 1384.0, 3)}), flush=True)


if __name__ == "__main__":
    main()
