#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native Chess-AI MLP evaluator.

Metric (BASELINE.json): MLP board evals/sec (+ train samples/sec). Workload at N=1 = BASELINE.json configs[1]:
  batched forward eval of the default MLP {64,144,225,100,1} (weights = the reference's network_dump/latest.txt)
  on 2^20 synthetic random legal boards, batch 4096 -> 256 launches per step, fp64 reference-precision path.
A "step" = one pass over the 2^20 boards. N>1 = weak scaling: every rank evaluates its own 2^20 boards
(boards are independent; no data-path collective), value = all boards of all ranks / max-over-ranks device time.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one JSON line on stdout)
  python bench.py --impl reference [...]                          the reference's own CPU code on the host cores

Keys beyond the base contract:
  roofline        fp64 tensor pipe (DMMA) of mlp_forward_kernel; frac = overlapped launches, frac_single_launch = one launch alone
  cpu_baseline    the reference's own code on the host cores (-O2 and as-shipped -O0 builds; configs[0] timed as stated)
  e2e             host buffers through cab_eval_host: H2D + kernels + D2H per step
  train           minibatch trainer resident on the device (cab_train_epoch_dev), weak scaling, NCCL all-reduce per step,
                  replicas_bit_identical = sha256 of every rank's weights after the timed steps
  dp_epoch        configs[3]: 10 x 2^20 synthetic (board, target) pairs split over the ranks (strong scaling)
  selfplay        configs[2]: 512 games per GPU x 800 playouts per move, MCTS resident on the device
  wide_bf16       configs[4]: 64-2048x4-1 on the bf16 tcgen05 path, forward and forward+backward (N=1 only)
  allreduce_514KB NCCL all-reduce latency at the trainer's message size (N>1)
"""
import argparse
import ctypes as C
import hashlib
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

N_BOARDS = 1 << 20          # per GPU per step
BATCH = 4096
N_INPUT_SETS = 4            # rotating input sets: 4 x 64 MiB = 256 MiB > 126 MB L2
FLOP_PER_EVAL = 128432      # 2 * 64 216 MACs; activations not counted
N_WEIGHTS = 64216
DEFAULT_DIMS = [64, 144, 225, 100, 1]
WIDE_DIMS = [64, 2048, 2048, 2048, 2048, 1]
GOLDEN = os.path.join(ROOT, "tests", "golden")
CPU_SAMPLE_BOARDS = 32768
B_TRAIN = 65536
EPOCH_SETS = 10             # configs[3]: 10 x 2^20 = 10 485 760 pairs in total


def latest_weights():
    return np.fromfile(os.path.join(GOLDEN, "latest_weights.f64"), dtype="<f8")


def _gen_boards(job):
    synth = importlib.import_module("chess-ai_b200.synth")
    return synth.synthetic_boards(job[1], start_chunk=job[0])


def make_board_sets(first_chunk, n_sets, workers):
    """n_sets x 2^20 synthetic boards (chunks first_chunk ...) by `workers` host processes; forked before CUDA exists."""
    per = 2 * 65536
    jobs = [(first_chunk + i // 65536, per) for i in range(0, n_sets * N_BOARDS, per)]
    if workers <= 1:
        return np.concatenate([_gen_boards(j) for j in jobs])
    with ProcessPoolExecutor(max_workers=workers) as ex:
        return np.concatenate(list(ex.map(_gen_boards, jobs)))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_rate(codes, threads, reps, as_shipped=False):
    """Reference CPU path (oracle/_ref: the reference's own predictPosition) or, failing that, the C port."""
    from oracle_bind import REF_SO, REF_SO_O0, PortOracle, RefOracle
    w = latest_weights()
    path = REF_SO_O0 if as_shipped else REF_SO
    if os.path.exists(path):
        ro = RefOracle(path)
        net = ro.net_from_weights(DEFAULT_DIMS, w)
        secs, out = net.time_predict(codes, threads=threads, reps=reps)
        return len(codes) * reps / secs, "reference", threads, out
    if as_shipped:
        raise FileNotFoundError(path)
    po = PortOracle()
    net = po.net_from_weights(DEFAULT_DIMS, w)
    secs, out = net.time_predict(codes)
    return len(codes) / secs, "port", 1, out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cab = importlib.import_module("chess-ai_b200")
    cores = os.cpu_count() or 1
    codes = cab.synthetic_boards(CPU_SAMPLE_BOARDS)
    reps = 2
    kind = "reference"
    for _ in range(args.warmup):
        cpu_reference_rate(codes[:4096], cores, 1)
    t0 = time.perf_counter()
    rates = []
    for _ in range(args.steps):
        rate, kind, used, _ = cpu_reference_rate(codes, cores, reps)
        rates.append(rate)
    dt = time.perf_counter() - t0
    value = float(np.mean(rates))
    sample = "%d synthetic boards x %d passes per step, all %d host threads, predictPosition on pre-built boards" % (
        CPU_SAMPLE_BOARDS, reps, used)
    line = {
        "impl": "reference", "metric": "MLP board evals/sec", "value": value, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / max(args.steps, 1) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),      # the SAME config as our arm; each step times a bounded sample of it:
        "reference_sample": {"boards_per_step": CPU_SAMPLE_BOARDS * reps, "what": sample,
                             "note": "config names the workload both arms are rated on; this arm evaluates a bounded sample of it per step "
                                     "(evals/s is a rate, so the sample size does not enter the comparison)"},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": used, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(n_gpus):
    return {"workload": "configs[1]: default MLP 64-144-225-100-1 (latest.txt weights), forward eval of 2^20 synthetic "
                        "random legal boards per GPU in launches of 4096, fp64 reference-precision path",
            "boards_per_gpu_per_step": N_BOARDS, "batch": BATCH, "launches_per_step": N_BOARDS // BATCH,
            "parallelism": "independent boards per GPU x%d (no collective)" % n_gpus,
            "l2": "%d rotating 64 MiB input sets (256 MiB > 126 MB L2)" % N_INPUT_SETS}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=int(os.environ.get("CAB_BENCH_STREAMS", "4")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training legs (train, dp_epoch)")
    ap.add_argument("--no-selfplay", action="store_true", help="skip the configs[2] self-play leg")
    ap.add_argument("--no-wide", action="store_true", help="skip the configs[4] bf16 leg")
    ap.add_argument("--selfplay-moves", type=int, default=20)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference_arm(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # ---- synthetic inputs, generated by host worker processes BEFORE CUDA is initialised in this process ------------
    # per rank: N_INPUT_SETS eval sets; the training legs reuse them and, for the rank's share of configs[3]'s
    # EPOCH_SETS x 2^20 pairs, as many more as the share needs
    epoch_local = (EPOCH_SETS * N_BOARDS // world) // B_TRAIN * B_TRAIN
    n_sets = N_INPUT_SETS if args.no_train else max(N_INPUT_SETS, -(-epoch_local // N_BOARDS))
    workers = max(1, (os.cpu_count() or 8) // world)
    host_all = make_board_sets(rank * 16 * (N_BOARDS // 65536), n_sets, workers)

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: chess-ai_b200 has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cab = importlib.import_module("chess-ai_b200")
    net = cab.Network.from_weights(DEFAULT_DIMS, latest_weights(), device=local_rank)
    warmup = max(args.warmup, 3)

    dev_all = torch.from_numpy(host_all).cuda()
    host_sets = [host_all[s * N_BOARDS:(s + 1) * N_BOARDS] for s in range(N_INPUT_SETS)]
    dev_sets = [dev_all[s * N_BOARDS:(s + 1) * N_BOARDS] for s in range(N_INPUT_SETS)]
    dev_out = torch.empty(N_BOARDS, dtype=torch.float64, device="cuda")
    streams = [torch.cuda.Stream() for _ in range(max(args.streams, 1))]
    stream_ptrs = [s.cuda_stream for s in streams]
    main_stream = torch.cuda.current_stream()

    def step_device(i):
        x = dev_sets[i % N_INPUT_SETS]
        fork = torch.cuda.Event()
        fork.record(main_stream)
        for s in streams:
            s.wait_event(fork)
        net.eval_dev_batches(x.data_ptr(), N_BOARDS, BATCH, dev_out.data_ptr(), stream_ptrs)
        for s in streams:
            ev = torch.cuda.Event()
            ev.record(s)
            main_stream.wait_event(ev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    for i in range(warmup):
        step_device(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    launches0 = net.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(main_stream)
    for i in range(args.steps):
        step_device(i)
    e1.record(main_stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = net.launch_count - launches0
    sampler.stop()
    ms = max_over_ranks(ms)
    value = world * N_BOARDS * args.steps / (ms * 1e-3)
    # the roofline's denominator, measured right behind the timed region (same clocks and temperature)
    peak = cab.measure_pipe_peak("dmma", local_rank) if rank == 0 else None

    # parity spot check (not timed): 2048 golden boards vs the reference's recorded outputs
    golden = np.load(os.path.join(GOLDEN, "golden.npz"))
    parity = None
    if rank == 0:
        y = net.predict_batch(golden["syn_codes"])
        want = golden["syn_forward_latest"]
        parity = float(np.max(np.abs(y - want) / np.maximum(np.abs(want), 1e-12)))

    # ---- per-launch duration of the dominant kernel on ONE stream (CUDA events on the launching stream) -------------
    s0 = streams[0]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    x = dev_sets[0]
    torch.cuda.synchronize()
    n_l = 64
    with torch.cuda.stream(s0):
        ev[0].record(s0)
        net.eval_dev_batches(x.data_ptr(), n_l * BATCH, BATCH, dev_out.data_ptr(), stream_ptrs[:1])
        ev[1].record(s0)
    torch.cuda.synchronize()
    launch_us_serial = ev[0].elapsed_time(ev[1]) * 1e3 / n_l

    # ---- the optional reduced-precision path on the same workload (reported beside the headline, never as `value`) ---
    f16_path = None
    if rank == 0:
        try:
            y16 = net.predict_batch_f16(golden["syn_codes"])
            y64 = golden["syn_forward_latest"]
            strong = np.abs(y64) >= 0.15
            agree = float((np.sign(y16[strong]) == np.sign(y64[strong])).mean())

            f16_streams = [torch.cuda.Stream() for _ in range(8)]     # a 4096-board launch of this kernel is ~5 us: 8 in flight
            f16_ptrs = [s.cuda_stream for s in f16_streams]

            def f16_step(i, batch):
                x = dev_sets[i % len(dev_sets)]
                net.eval_f16_dev_batches(x.data_ptr(), N_BOARDS, batch, dev_out.data_ptr(), f16_ptrs)
            res = {}
            for name, batch in (("batch_%d" % BATCH, BATCH), ("one_launch_per_step", N_BOARDS)):
                for i in range(len(dev_sets)):          # every input set once: its launch graph is captured here, not in the timed loop
                    f16_step(i, batch)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for i in range(5):
                    f16_step(i, batch)
                torch.cuda.synchronize()
                res[name] = N_BOARDS * 5 / (time.perf_counter() - t0)
            f16_path = {"api": "cab_eval_f16_dev (fused tcgen05 kernel, fp16 operands, fp32 accumulate, activations in tensor memory)",
                        "evals_per_s": res, "unit": "evals/s per GPU, inputs resident in HBM, wall clock around 5 steps", "streams": 8,
                        "launching": "the launches of one call are replayed as a CUDA graph (captured once per buffers / sizes / streams)",
                        "sign_agreement_vs_fp64_on_golden_boards": agree,
                        "parity": "reduced precision: no value parity; move choice made identical to fp64 by cab_best_move_host (tests/test_tc_gpu.py)"}
        except Exception as exc:  # the path is optional: never take the headline down with it
            f16_path = {"error": str(exc)}

    # ---- e2e: host (pinned) buffers through the C-ABI host entry point, H2D and D2H inside the timed region --------
    pin_in = [torch.from_numpy(h).pin_memory() for h in host_sets[:2]]
    pin_out = torch.empty(N_BOARDS, dtype=torch.float64).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    for i in range(2):
        net.eval_host_ptr(pin_in[i % 2].data_ptr(), N_BOARDS, pin_out.data_ptr())
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        net.eval_host_ptr(pin_in[i % 2].data_ptr(), N_BOARDS, pin_out.data_ptr())
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * N_BOARDS * e2e_steps / e2e_s
    del pin_in

    # ---- NCCL all-reduce latency at the trainer's message size (64 220 doubles = 514 KB) --------------
    allreduce = None
    if world > 1:
        buf = torch.zeros(N_WEIGHTS + 4, dtype=torch.float64, device="cuda")
        for _ in range(20):
            dist.all_reduce(buf)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(200):
            dist.all_reduce(buf)
        a1.record()
        torch.cuda.synchronize()
        allreduce = {"us": max_over_ranks(a0.elapsed_time(a1) * 1e3 / 200), "bytes": (N_WEIGHTS + 4) * 8, "ranks": world,
                     "how": "200 back-to-back ncclAllReduce(sum, fp64) on one stream, CUDA events, max over ranks"}

    # ---- training (second half of BASELINE.json's metric) ------------------------------------------------------------
    # the minibatch extension of Optimizer::optimize with everything on the device (cab_train_epoch_dev): teacher targets =
    # latest.txt outputs + N(0, 0.05) clipped to [-4, 4], student = U(-1,1) x 0.1 with the same seed on every rank,
    # lambda-reset + dropout rule of train.cpp:54-57 active, NCCL all-reduce of the gradient inside the library.
    train = dp_epoch = None
    if not args.no_train:
        dpm = importlib.import_module("chess-ai_b200.dp")
        n_local = dev_all.shape[0]
        targets = torch.empty(n_local, dtype=torch.float64, device="cuda")
        net.eval_dev_ptr(dev_all.data_ptr(), n_local, targets.data_ptr(), main_stream.cuda_stream)
        torch.cuda.synchronize()
        gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
        targets = (targets + 0.05 * torch.randn(n_local, dtype=torch.float64, device="cuda", generator=gen)).clamp_(-4, 4)
        def new_student():
            s_net = cab.Network.randomNetwork(DEFAULT_DIMS, device=local_rank, seed=1234)   # same seed: identical replicas
            s_net.set_weights(s_net.get_weights() * 0.1)
            return s_net
        student = new_student()
        # the collective of the data-parallel step: fused into the update kernel over NVLink peer memory (cab_dp_peer_*);
        # NCCL all-reduces (cab_dp_init) if the ranks cannot map each other's memory, and as a second timing beside it
        par, collective, nccl_note = None, "none", None
        if world > 1:
            mode = os.environ.get("CAB_BENCH_COLLECTIVE", "peer")
            try:
                par = dpm.DataParallel(cab, student, peer=(mode == "peer"))
                collective = ("gradient sum over peer-mapped buffers fused into the update kernel + 8-byte peer stores for E' (no NCCL in the step)"
                              if mode == "peer" else "ncclAllReduce(64220 fp64) + 1 scalar per step")
            except Exception as exc:             # every rank fails alike (same box): fall back together
                nccl_note = "peer mode unavailable (%s): NCCL" % exc
                student = new_student()
                par = dpm.DataParallel(cab, student)
                collective = "ncclAllReduce(64220 fp64) + 1 scalar per step"

        def run_epoch(first, n, lam, seed):
            """n resident cases from `first` on as minibatches of B_TRAIN; device time of the whole call, max over ranks."""
            barrier()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = student.launch_count
            t0.record(streams[0])
            lam, st = cab.train_epoch_dev(student, dev_all.data_ptr() + first * 64, targets.data_ptr() + first * 8, n, B_TRAIN, lam,
                                          reset_rule=True, dropout_seed=seed, stream=stream_ptrs[0])
            t1.record(streams[0])
            torch.cuda.synchronize()
            return lam, st, max_over_ranks(t0.elapsed_time(t1) * 1e-3), int(student.launch_count - l0)

        def digest_check():
            h = hashlib.sha256(student.get_weights().tobytes()).hexdigest()
            if world == 1:
                return True, h
            hs = [None] * world
            dist.all_gather_object(hs, h)
            return all(x == hs[0] for x in hs), h

        # weak scaling: every rank trains 64 steps of 65 536 of ITS boards (its four eval sets); 16 untimed steps first so
        # that lambda has left its start value and the timed steps see the steady accept / reject mix
        t_steps = N_INPUT_SETS * N_BOARDS // B_TRAIN
        lam, _, _, _ = run_epoch(0, 16 * B_TRAIN, 0.5, 0x5EED)
        lam, st, secs, n_launch = run_epoch(0, t_steps * B_TRAIN, lam, 0x5EED + 1)
        same, digest = digest_check()
        train = {"value": world * B_TRAIN * t_steps / secs, "unit": "samples/s", "batch_per_gpu": B_TRAIN, "steps": t_steps,
                 "ms_per_step": secs / t_steps * 1e3, "accepted_steps": st["accepted"], "resets": st["resets"],
                 "E_first": st["first_err"], "E_last": st["last_new_err"] if st["last_accepted"] else st["last_err"], "lambda_end": lam,
                 "flop_per_sample": 8 * N_WEIGHTS, "flop_note": "8P per sample (fwd 2P + dX 2P + dW 2P + re-fwd 2P); the update is per step",
                 "tflops": 8 * N_WEIGHTS * world * B_TRAIN * t_steps / secs * 1e-12,
                 "gpu_launches": n_launch, "api": "cab_train_epoch_dev: decisions, rollback and reset+dropout on the device, one wait at the end",
                 "collective": collective, "collective_note": nccl_note,
                 "replicas_bit_identical": same, "weights_sha256": digest[:16], "scaling": "weak",
                 "timing": "CUDA events around the whole call on the launching stream, max over ranks"}

        # configs[3]: EPOCH_SETS x 2^20 pairs in total, rank r owns total / N of them (strong scaling), global batch N x 65 536
        lam, st, secs, n_launch = run_epoch(0, epoch_local, 2.0, 0x5EED + 2)
        same, digest = digest_check()
        dp_epoch = {"value": world * epoch_local / secs, "unit": "samples/s", "pairs_total": world * epoch_local, "pairs_per_gpu": epoch_local,
                    "batch_per_gpu": B_TRAIN, "steps": st["steps"], "seconds": secs, "accepted_steps": st["accepted"], "resets": st["resets"],
                    "E_first": st["first_err"], "E_last": st["last_new_err"] if st["last_accepted"] else st["last_err"], "lambda_end": lam,
                    "replicas_bit_identical": same, "weights_sha256": digest[:16], "scaling": "strong", "gpu_launches": n_launch,
                    "config": "configs[3]: data-parallel training of the default MLP on 10 x 2^20 synthetic pairs, dropout rule active, lambda0 = 2.0"}
        if par is not None:
            par.shutdown()
        if world > 1 and par is not None and par.peer:      # the same weak-scaling steps with NCCL all-reduces, for comparison
            try:
                student = new_student()
                par = dpm.DataParallel(cab, student)
                lam, _, _, _ = run_epoch(0, 16 * B_TRAIN, 0.5, 0x5EED)
                lam, st, secs, _ = run_epoch(0, t_steps * B_TRAIN, lam, 0x5EED + 1)
                same, _ = digest_check()
                train["with_nccl_allreduce"] = {"value": world * B_TRAIN * t_steps / secs, "ms_per_step": secs / t_steps * 1e3,
                                                "accepted_steps": st["accepted"], "replicas_bit_identical": same}
                par.shutdown()
            except Exception as exc:
                train["with_nccl_allreduce"] = {"error": str(exc)}
        del targets, student

    # ---- configs[2]: MCTS self-play resident on the device, independent games per GPU (no collective) ------------------
    selfplay = None
    if not args.no_selfplay:
        sp = cab.SelfPlay(net, games=512, playouts_per_move=800, roots=8, in_flight=16, max_moves=args.selfplay_moves, network_priors=1,
                          noise=1, save_ratio=0.1, seed=1234 + rank)
        barrier()
        st = sp.run()
        secs = max_over_ranks(st["seconds"])
        tot = {k: sum_over_ranks(st[k]) for k in ("playouts", "boards_evaluated", "moves_played", "expansions", "cases_kept")}
        selfplay = {"value": tot["playouts"] / secs, "unit": "playouts/s", "positions_evaluated_per_s": tot["boards_evaluated"] / secs,
                    "games_per_gpu": 512, "playouts_per_move": 800, "moves_per_game": args.selfplay_moves, "roots": 8, "in_flight_per_tree": 16,
                    "network_priors": 1, "noise": 1, "seconds": secs, "moves_played": tot["moves_played"], "expansions": tot["expansions"],
                    "cases_kept": tot["cases_kept"], "evaluator_busy_rank0": st["evaluator_seconds"] / st["seconds"],
                    "rounds_rank0": st["rounds"], "arena_full_rank0": st["arena_full"], "eval_full_rank0": st["eval_full"], "scaling": "weak",
                    "config": "configs[2]: 512 concurrent games per GPU, 800 playouts per move, virtual-loss leaf batching; trees, "
                              "move generation, expansion, priors and game rules on the device (cab_mcts_run), fp64 evaluator",
                    "timing": "wall clock inside cab_mcts_run (it ends with a device synchronise), max over ranks"}
        sp.close()
        # the same games with the fused fp16 tcgen05 evaluator behind the search (reduced precision: not the reference's search)
        try:
            sp = cab.SelfPlay(net, games=512, playouts_per_move=800, roots=8, in_flight=16, max_moves=args.selfplay_moves, network_priors=1,
                              noise=1, save_ratio=0.1, seed=1234 + rank, precision=1)
            barrier()
            st = sp.run()
            secs = max_over_ranks(st["seconds"])
            selfplay["f16_evaluator"] = {"value": sum_over_ranks(st["playouts"]) / secs, "unit": "playouts/s",
                                         "positions_evaluated_per_s": sum_over_ranks(st["boards_evaluated"]) / secs, "seconds": secs,
                                         "evaluator_busy_rank0": st["evaluator_seconds"] / st["seconds"],
                                         "note": "cab_mcts_params.precision = 1: priors and leaf values from cab_eval_f16_dev"}
            sp.close()
        except Exception as exc:
            selfplay["f16_evaluator"] = {"error": str(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- K0 encoder (Network::loadBoard's integer stage): HBM-bound byte kernel, 192 B in + 64 B out per board ------------
    encoder = None
    try:
        hbm_peak, hbm_src = 6539.2, "fallback (the pool's recorded copy bandwidth)"
        ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(ppath):
            hbm_peak, hbm_src = float(json.load(open(ppath)).get("hbm_gbs", hbm_peak)), "MEASURED_PEAKS.json hbm_gbs (copy, read+write bytes)"
        n_enc = 1 << 21                                     # 384 MiB of records + 128 MiB of codes: larger than L2
        recs = torch.randint(0, 7, (n_enc, 64, 3), dtype=torch.uint8, device="cuda")
        enc_out = torch.empty(n_enc, 64, dtype=torch.int8, device="cuda")
        L = cab.lib()
        for _ in range(3):
            cab._ck(L.cab_encode_dev(recs.data_ptr(), n_enc, enc_out.data_ptr(), stream_ptrs[0]))
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(streams[0])
        for _ in range(10):
            cab._ck(L.cab_encode_dev(recs.data_ptr(), n_enc, enc_out.data_ptr(), stream_ptrs[0]))
        k1.record(streams[0])
        torch.cuda.synchronize()
        enc_s = k0.elapsed_time(k1) * 1e-3 / 10
        encoder = {"kernel": "encode_kernel", "bound": "hbm", "boards_per_launch": n_enc, "bytes_per_board": 256,
                   "launch_us": enc_s * 1e6, "achieved": 256 * n_enc / enc_s * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                   "frac": 256 * n_enc / enc_s * 1e-9 / hbm_peak, "peak_source": hbm_src, "boards_per_s": n_enc / enc_s}
        del recs, enc_out
    except Exception as exc:
        encoder = {"error": str(exc)}

    # ---- configs[4]: widened MLP on the bf16 tcgen05 path (N = 1 runs only) ---------------------------------------------
    wide = None
    if not args.no_wide and world == 1:
        try:
            peaks = {"bf16_tflops": 1627.0, "bf16_tflops_sustained": 1384.0}
            ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
            if os.path.exists(ppath):
                peaks.update({k: v for k, v in json.load(open(ppath)).items() if k in peaks})
            p_gemm = sum(k * n for k, n in zip(WIDE_DIMS[:-1], WIDE_DIMS[1:]))
            flop_fwd = 2 * p_gemm
            flop_fb = 2 * p_gemm + 2 * sum(k * n for k, n in list(zip(WIDE_DIMS[:-1], WIDE_DIMS[1:]))[1:]) + 2 * p_gemm
            wnet = cab.Network.randomNetwork(WIDE_DIMS, device=local_rank, seed=1)
            L = cab.lib()
            wcodes = dev_sets[2][:65536]
            wtargets = torch.randn(65536, dtype=torch.float64, device="cuda").clamp_(-3.9, 3.9)
            wstream, sptr = streams[0], stream_ptrs[0]     # a real stream: a null stream would select the library's own one

            def timed(fn, reps):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                w0.record(wstream)
                for _ in range(reps):
                    fn()
                w1.record(wstream)
                torch.cuda.synchronize()
                return w0.elapsed_time(w1) * 1e-3 / reps
            wide = {"dims": WIDE_DIMS, "peak_burst": peaks["bf16_tflops"], "peak_sustained": peaks["bf16_tflops_sustained"],
                    "flop_per_board_forward": flop_fwd, "flop_per_board_forward_backward": flop_fb, "runs": []}
            for wb in (16384, 65536):
                t_f = timed(lambda: wnet.eval_bf16_dev_ptr(wcodes.data_ptr(), wb, dev_out.data_ptr(), sptr), 20)
                t_fb = timed(lambda: cab._ck(L.cab_grad_bf16_dev(wnet._h, wcodes.data_ptr(), wtargets.data_ptr(), wb, sptr)), 10)
                tf_f, tf_fb = flop_fwd * wb / t_f * 1e-12, flop_fb * wb / t_fb * 1e-12
                wide["runs"].append({"batch": wb, "forward_tflops": tf_f, "forward_frac_sustained": tf_f / peaks["bf16_tflops_sustained"],
                                     "forward_frac_burst": tf_f / peaks["bf16_tflops"], "forward_backward_tflops": tf_fb,
                                     "forward_backward_frac_sustained": tf_fb / peaks["bf16_tflops_sustained"],
                                     "forward_backward_frac_burst": tf_fb / peaks["bf16_tflops"]})
            del wnet
        except Exception as exc:
            wide = {"error": str(exc)}

    # ---- roofline of the dominant kernel (mlp_forward_kernel) on the fp64 tensor pipe -------------------------------
    n_fwd_launches = args.steps * (N_BOARDS // BATCH)
    avg_launch_s = (ms * 1e-3) / n_fwd_launches          # launches overlap on the streams: region time / launches
    achieved = FLOP_PER_EVAL * BATCH / avg_launch_s * 1e-12
    traffic = traffic_cold = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("mlp_forward_kernel_dram_bytes_per_launch_steady_state")
            traffic_cold = tj.get("mlp_forward_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "tensor", "pipe": "fp64 tensor pipe (DMMA.8x8x4)", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak,
                "frac_single_launch": FLOP_PER_EVAL * BATCH / (launch_us_serial * 1e-6) * 1e-12 / peak,
                "traffic": traffic, "traffic_note": "steady state (weights L2-resident, ncu --cache-control none inside the bench loop)",
                "traffic_cold_cache": traffic_cold, "algorithmic_bytes_per_launch": 72 * BATCH,
                "peak_source": "measured in this run with cab_measure_pipe_peak (best of 10, CUDA events); "
                               "MEASURED_PEAKS.json carries no fp64 figure",
                "flop_per_unit": FLOP_PER_EVAL, "units_per_launch": BATCH,
                "avg_launch_us_concurrent": avg_launch_s * 1e6, "launch_us_single_stream": launch_us_serial}

    cpu_baseline = None
    if not args.no_cpu_baseline and args.gpus == 1:
        cores = os.cpu_count() or 1
        codes = host_sets[0][:CPU_SAMPLE_BOARDS]
        rate, kind, used, out = cpu_reference_rate(codes, cores, 4)
        y = net.predict_batch(codes)
        agree = float(np.max(np.abs(y - out) / np.maximum(np.abs(out), 1e-12)))
        cpu_baseline = {"value": rate, "unit": "evals/s", "cores": used, "kind": kind,
                        "sample": "%d of the same synthetic boards x 4 passes, one Network::clone() per thread, g++ -O2" % CPU_SAMPLE_BOARDS,
                        "max_rel_diff_vs_gpu": agree}
        try:  # configs[0] also asks for the single-thread rate
            r1, _, _, _ = cpu_reference_rate(codes[:4096], 1, 1)
            cpu_baseline["value_1_thread"] = r1
        except Exception as exc:
            cpu_baseline["value_1_thread_error"] = str(exc)
        try:  # the reference's as-shipped flags: its CMakeLists.txt sets no optimisation level
            r0, _, u0, _ = cpu_reference_rate(codes[:8192], cores, 1, as_shipped=True)
            cpu_baseline["value_as_shipped_flags"] = r0
            cpu_baseline["as_shipped_sample"] = "8192 boards x 1 pass, %d threads, the same translation units at -O0 -DDEBUG (CMakeLists.txt:1-49)" % u0
        except Exception as exc:
            cpu_baseline["value_as_shipped_flags_error"] = str(exc)
        try:  # Optimizer::optimize is inherently sequential per network: 1 thread
            from oracle_bind import RefOracle
            if RefOracle.available():
                ro = RefOracle()
                rnet = ro.net_from_weights(DEFAULT_DIMS, latest_weights())
                n_t = 2048
                secs, _ = rnet.time_train(codes[:n_t], np.zeros(n_t), 2.0)
                cpu_baseline["train_value"] = n_t / secs
                cpu_baseline["train_unit"] = "samples/s"
                cpu_baseline["train_sample"] = "%d Optimizer::optimize steps on synthetic boards, 1 thread" % n_t
                # configs[0] exactly as stated: forward over the 66 shipped cases + ONE optimize epoch over them in file order
                cc, ct = golden["case_codes"], golden["case_targets"]
                rnet = ro.net_from_weights(DEFAULT_DIMS, latest_weights())
                fsecs, _ = rnet.time_predict(cc, threads=1, reps=20)
                esecs, lam_ref = rnet.time_train(cc, ct, 2.0)
                gnet = cab.Network.from_weights(DEFAULT_DIMS, latest_weights(), device=local_rank)
                gnet.predict_batch(cc)
                t0 = time.perf_counter()
                for _ in range(20):
                    gnet.predict_batch(cc)
                gf = (time.perf_counter() - t0) / 20
                cab.Optimizer.train_sequence(gnet.clone(), cc, ct, 2.0)
                t0 = time.perf_counter()
                lam_gpu = cab.Optimizer.train_sequence(gnet, cc, ct, 2.0)[0]
                ge = time.perf_counter() - t0
                cpu_baseline["configs0"] = {
                    "what": "configs[0]: latest.txt, forward over the 66 training_cases boards + one Optimizer::optimize epoch in file order, lambda0 = 2.0",
                    "cpu_forward_evals_per_s_1_thread": 66 * 20 / fsecs, "cpu_epoch_seconds": esecs, "cpu_epoch_samples_per_s": 66 / esecs,
                    "cpu_lambda_end": lam_ref,
                    "gpu_forward_evals_per_s_host_call": 66 / gf, "gpu_epoch_seconds": ge, "gpu_epoch_samples_per_s": 66 / ge,
                    "gpu_lambda_end": lam_gpu, "lambda_identical": bool(lam_ref == lam_gpu),
                    "note": "B = 1 steps are strictly sequential (each step reads the previous step's weights): the GPU runs them as one "
                            "persistent CTA (train_sequence_kernel) and is latency-bound here by construction"}
        except Exception as exc:  # the eval baseline stands on its own
            cpu_baseline["train_error"] = str(exc)

    line = {
        "metric": "MLP board evals/sec", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(world),
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": N_BOARDS * 64, "d2h_bytes_per_step": N_BOARDS * 8,
                "steps": e2e_steps, "api": "cab_eval_host (pinned host buffers)"},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": sampler.summary(),
        "parity_max_rel_err_vs_reference_golden": parity, "streams": len(streams), "train": train, "dp_epoch": dp_epoch,
        "selfplay": selfplay, "wide_bf16": wide, "encoder": encoder, "allreduce_514KB": allreduce,
        "precision_paths": {"f16_tcgen05": f16_path},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
