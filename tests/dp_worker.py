"""Worker for tests/test_dp.py, launched with torch.distributed.run (2 ranks).

  --mode gloo : CPU-only host logic — shard cover, NCCL-id broadcast plumbing, and the data-parallel reduction rule
                (sum of per-rank (sum dW-equivalent, sum E, count) gives the same mean and decision on every rank),
                with the C oracle standing in for the device.
  --mode nccl : two GPUs — DataParallel.train_batch on shards must equal single-GPU training on the whole batch and
                leave bit-identical weights on both ranks.
  --mode peer : the same with the reduction over NVLink peer memory fused into the update kernel (cab_dp_peer_*), plus an
                epoch of resident minibatches (cab_train_epoch_dev) against the same epoch on one GPU.
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
DIMS = [64, 144, 225, 100, 1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", required=True)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    cab = importlib.import_module("chess-ai_b200")
    dp = importlib.import_module("chess-ai_b200.dp")
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))
    codes = g["syn_codes"][:1001]
    rng = np.random.default_rng(5)
    targets = rng.uniform(-2, 2, len(codes))
    w0 = np.random.default_rng(11).uniform(-0.3, 0.3, 64216)
    res = {"rank": rank}
    if a.mode == "gloo":
        dist.init_process_group("gloo")
        from oracle_bind import PortOracle
        po = PortOracle()
        lo, hi = dp.shard_range(len(codes), rank, world)
        spans = [None] * world
        dist.all_gather_object(spans, (lo, hi))
        res["spans"] = spans
        ident = bytes(range(128)) if rank == 0 else bytes(128)
        res["ident_ok"] = dp.broadcast_bytes(ident, src=0) == bytes(range(128))
        # reduction rule: per-rank sums of E over the shard, all-reduced, equal the whole-batch mean on every rank
        net = po.net_from_weights(DIMS, w0)
        y = net.predict(codes[lo:hi])
        part = torch.tensor([np.sum((targets[lo:hi] - y) ** 2 / 2.0), hi - lo], dtype=torch.float64)
        dist.all_reduce(part)
        whole = np.sum((targets - net.predict(codes)) ** 2 / 2.0) / len(codes)
        res["mean_err"] = float(part[0] / part[1])
        res["whole_err"] = float(whole)
    else:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
        net = cab.Network.from_weights(DIMS, w0, device=rank)
        par = dp.DataParallel(cab, net, peer=(a.mode == "peer"))
        lo, hi = dp.shard_range(len(codes), rank, world)
        lam, trace = 0.5, []
        for step in range(4):
            lam, acc, e0, e1 = par.train_batch(codes[lo:hi], targets[lo:hi], lam)
            trace.append([lam, acc, e0, e1])
        w = net.get_weights()
        res["trace"] = trace
        res["w_sha"] = __import__("hashlib").sha256(w.tobytes()).hexdigest()
        if rank == 0:
            single = cab.Network.from_weights(DIMS, w0, device=0)
            lam_s, trace_s = 0.5, []
            for step in range(4):
                lam_s, acc, e0, e1 = cab.Optimizer.train_batch(single, codes, targets, lam_s)
                trace_s.append([lam_s, acc, e0, e1])
            res["trace_single"] = trace_s
            res["max_abs_w_diff_vs_single"] = float(np.abs(w - single.get_weights()).max())
        if a.mode == "peer":   # resident epoch: 6 minibatches of 100 per rank (+ a short one), decisions on the device, one wait
            big = np.ascontiguousarray(np.tile(g["syn_codes"], (2, 1))[:1300])
            tb = np.random.default_rng(9).uniform(-2, 2, len(big))
            per = len(big) // world
            mine_c = torch.from_numpy(np.ascontiguousarray(big[rank * per:(rank + 1) * per])).cuda()
            mine_t = torch.from_numpy(np.ascontiguousarray(tb[rank * per:(rank + 1) * per])).cuda()
            s = torch.cuda.Stream()
            lam_e, st, tr = cab.train_epoch_dev(net, mine_c.data_ptr(), mine_t.data_ptr(), per, 100, 0.5, trace=True, stream=s.cuda_stream)
            res["epoch"] = {"lam": lam_e, "stats": st, "trace": tr.tolist()}
            res["w_sha_epoch"] = __import__("hashlib").sha256(net.get_weights().tobytes()).hexdigest()
            if rank == 0:      # the same global minibatches on one GPU: rank 0's 100 boards + rank 1's 100 boards per step
                single2 = cab.Network.from_weights(DIMS, w, device=0)
                order = np.concatenate([np.concatenate([np.arange(r * per + b0, min(r * per + b0 + 100, (r + 1) * per)) for r in range(world)])
                                        for b0 in range(0, per, 100)])
                oc = torch.from_numpy(np.ascontiguousarray(big[order])).cuda()
                ot = torch.from_numpy(np.ascontiguousarray(tb[order])).cuda()
                lam_1, st1, tr1 = cab.train_epoch_dev(single2, oc.data_ptr(), ot.data_ptr(), len(order), 100 * world, 0.5, trace=True, stream=s.cuda_stream)
                res["epoch_single"] = {"lam": lam_1, "stats": st1, "trace": tr1.tolist()}
                res["epoch_max_abs_w_diff_vs_single"] = float(np.abs(net.get_weights() - single2.get_weights()).max())
        par.shutdown()
    with open("%s.%d.json" % (a.out, rank), "w") as f:
        json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
