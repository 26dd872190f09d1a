"""BASELINE.json configs[3]: data-parallel training over 10 M synthetic (board, target) pairs, teacher = latest.txt
(+ N(0, 0.05), clipped to [-4, 4]), lambda reset + dropout rule active, one NCCL all-reduce of the gradient per step
inside the library. Run alone (N = 1) or under torchrun (one process per GPU, 127.0.0.1 rendezvous):

    python scripts/dp_train_epoch.py [total_pairs] [batch_per_gpu]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/dp_train_epoch.py

Every rank keeps its shard of the pairs resident in HBM; a step takes the next batch_per_gpu pairs of every shard
(global batch = N x batch_per_gpu, weak scaling in the step, strong scaling in the epoch). Timed with a barrier and a
device synchronise on both sides, max over ranks. Rank 0 prints one JSON line."""
import ctypes as C
import importlib
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cab = importlib.import_module("chess-ai_b200")
synth = importlib.import_module("chess-ai_b200.synth")
dpm = importlib.import_module("chess-ai_b200.dp")
DIMS = [64, 144, 225, 100, 1]


def _gen(args):
    seed, chunk0, n = args
    return synth.synthetic_boards(n, seed=seed, start_chunk=chunk0)


def make_boards(n, first_chunk, workers):
    per = synth.CHUNK * 2
    jobs = [(0xC0FFEE, first_chunk + i // synth.CHUNK, min(per, n - i)) for i in range(0, n, per)]
    with ProcessPoolExecutor(max_workers=workers) as ex:
        return np.concatenate(list(ex.map(_gen, jobs)))


def main():
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    steps = total // (B * world)
    n_local = steps * B
    chunks_per_rank = (n_local + synth.CHUNK - 1) // synth.CHUNK
    t_gen = time.perf_counter()                                               # host worker processes fork before CUDA exists
    codes_h = make_boards(n_local, rank * chunks_per_rank, max(1, (os.cpu_count() or 8) // world))
    t_gen = time.perf_counter() - t_gen
    torch.cuda.set_device(local)
    if world > 1:
        global dist
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    codes = torch.from_numpy(codes_h).cuda()
    teacher = cab.Network.from_weights(DIMS, np.fromfile(os.path.join(ROOT, "tests", "golden", "latest_weights.f64"), dtype="<f8"), device=local)
    targets = torch.empty(n_local, dtype=torch.float64, device="cuda")
    stream = torch.cuda.Stream()
    teacher.eval_dev_ptr(codes.data_ptr(), n_local, targets.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    gen = torch.Generator(device="cuda").manual_seed(99 + rank)
    targets = (targets + 0.05 * torch.randn(n_local, dtype=torch.float64, device="cuda", generator=gen)).clamp_(-4, 4)
    student = cab.Network.randomNetwork(DIMS, device=local, seed=1234)      # same seed on every rank: identical replicas
    student.set_weights(student.get_weights() * 0.1)
    par = dpm.DataParallel(cab, student) if world > 1 else None
    L = cab.lib()

    def step(i, lam):
        lam_c, acc, e0, e1 = C.c_double(lam), C.c_int(0), C.c_double(0), C.c_double(0)
        cab._ck(L.cab_train_batch_dev(student._h, codes.data_ptr() + i * B * 64, targets.data_ptr() + i * B * 8, B, C.byref(lam_c),
                                      15.0, 1.1, C.byref(acc), C.byref(e0), C.byref(e1), None))
        return lam_c.value, bool(acc.value), e0.value, e1.value

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lam, resets, accepted = 2.0, 0, 0
    for i in range(min(3, steps)):                                            # warm-up (not counted, weights restored below)
        step(i, lam)
    student.set_weights(cab.Network.randomNetwork(DIMS, device=local, seed=1234).get_weights() * 0.1)
    l0 = student.launch_count
    barrier()
    t0 = time.perf_counter()
    e_first = e_last = None
    for i in range(steps):
        lam, acc, e0, e1 = step(i, lam)
        accepted += int(acc)
        e_first = e0 if e_first is None else e_first
        e_last = e1 if acc else e0
        if lam < 1e-4 or lam > 10.0:                                          # train.cpp:54-57, identical on every rank
            lam = 2.0
            cab.Optimizer.dropout(student, 0.01, seed=0x5EED, offset=resets)
            resets += 1
    barrier()
    dt = time.perf_counter() - t0
    w = student.get_weights()
    digest = float(np.abs(w).sum())
    if world > 1:
        t = torch.tensor([dt, digest, -digest], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, replicas_equal = float(t[0]), bool(t[1] == -t[2])                 # max(d) == min(d): bit-identical replicas
    else:
        replicas_equal = True
    if rank == 0:
        print(json.dumps({"config": "configs[3]: DP training over synthetic pairs, lambda reset + dropout rule active",
                          "n_gpus": world, "pairs": steps * B * world, "batch_per_gpu": B, "steps": steps, "seconds": round(dt, 4),
                          "samples_per_s": round(steps * B * world / dt, 1), "ms_per_step": round(dt / steps * 1e3, 3),
                          "accepted_steps": accepted, "lambda_resets": resets, "E_first": e_first, "E_last": e_last,
                          "lambda_final": lam, "replicas_bit_identical": replicas_equal,
                          "gpu_launches": int(student.launch_count - l0), "board_generation_s": round(t_gen, 1),
                          "collective": "ncclAllReduce(64220 fp64) + 1 scalar per step" if world > 1 else "none"}), flush=True)
    if par is not None:
        par.shutdown()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
