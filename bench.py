#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native Chess-AI MLP evaluator.

Metric (BASELINE.json): MLP board evals/sec. Workload at N=1 = BASELINE.json configs[1]:
  batched forward eval of the default MLP {64,144,225,100,1} (weights = the reference's network_dump/latest.txt)
  on 2^20 synthetic random legal boards, batch 4096 -> 256 launches per step, fp64 reference-precision path.
A "step" = one pass over the 2^20 boards. N>1 = weak scaling: every rank evaluates its own 2^20 boards
(boards are independent; no data-path collective), value = all boards of all ranks / max-over-ranks device time.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one JSON line on stdout)
  python bench.py --impl reference [...]                          the reference's own CPU code on the host cores

Keys beyond the base contract:  roofline (fp64 tensor pipe, DMMA), cpu_baseline (reference on the host cores),
e2e (host buffers through cab_eval_host: H2D + kernels + D2H per step), clocks, gpu_launches.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

N_BOARDS = 1 << 20          # per GPU per step
BATCH = 4096
N_INPUT_SETS = 4            # rotating input sets: 4 x 64 MiB = 256 MiB > 126 MB L2
FLOP_PER_EVAL = 128432      # 2 * 64 216 MACs; activations not counted
DEFAULT_DIMS = [64, 144, 225, 100, 1]
GOLDEN = os.path.join(ROOT, "tests", "golden")
CPU_SAMPLE_BOARDS = 32768


def latest_weights():
    return np.fromfile(os.path.join(GOLDEN, "latest_weights.f64"), dtype="<f8")


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


def cpu_reference_rate(codes, threads, reps):
    """Reference CPU path (oracle/_ref: the reference's own predictPosition) or, failing that, the C port."""
    from oracle_bind import PortOracle, RefOracle
    w = latest_weights()
    if RefOracle.available():
        ro = RefOracle()
        net = ro.net_from_weights(DEFAULT_DIMS, w)
        secs, out = net.time_predict(codes, threads=threads, reps=reps)
        return len(codes) * reps / secs, "reference", threads, out
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
        "config": workload_config(args.gpus),
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
    ap.add_argument("--no-train", action="store_true", help="skip the training-throughput section")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: chess-ai_b200 has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cab = importlib.import_module("chess-ai_b200")
    net = cab.Network.from_weights(DEFAULT_DIMS, latest_weights(), device=local_rank)
    warmup = max(args.warmup, 3)

    # ---- synthetic inputs: N_INPUT_SETS distinct sets per rank, resident in HBM --------------------------------
    host_sets = [cab.synthetic_boards(N_BOARDS, start_chunk=(rank * N_INPUT_SETS + s) * (N_BOARDS // 65536))
                 for s in range(N_INPUT_SETS)]
    dev_sets = [torch.from_numpy(h).cuda() for h in host_sets]
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
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * N_BOARDS * args.steps / (ms * 1e-3)

    # parity spot check of what was just computed (not timed): last step's first 2048 outputs vs golden
    golden = np.load(os.path.join(GOLDEN, "golden.npz"))
    parity = None
    if rank == 0:
        y = net.predict_batch(golden["syn_codes"])
        want = golden["syn_forward_latest"]
        parity = float(np.max(np.abs(y - want) / np.maximum(np.abs(want), 1e-12)))

    # ---- single-stream per-launch duration of the dominant kernel (CUDA events on the launching stream) ---------
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
    # fused fp16 tcgen05 evaluator (cab_eval_f16_dev): the same 2^20 boards per step, as launches of BATCH over the same
    # streams, and as one launch per step; rank 0's GPU only.
    f16_path = None
    if rank == 0:
        try:
            y16 = net.predict_batch_f16(golden["syn_codes"])
            y64 = golden["syn_forward_latest"]
            strong = np.abs(y64) >= 0.15
            agree = float((np.sign(y16[strong]) == np.sign(y64[strong])).mean())

            def f16_step(i, batch):
                x = dev_sets[i % len(dev_sets)]
                net.eval_f16_dev_batches(x.data_ptr(), N_BOARDS, batch, dev_out.data_ptr(), stream_ptrs)
            res = {}
            for name, batch in (("batch_%d" % BATCH, BATCH), ("one_launch_per_step", N_BOARDS)):
                for i in range(2):
                    f16_step(i, batch)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for i in range(5):
                    f16_step(i, batch)
                torch.cuda.synchronize()
                res[name] = N_BOARDS * 5 / (time.perf_counter() - t0)
            f16_path = {"api": "cab_eval_f16_dev (fused tcgen05 kernel, fp16 operands, fp32 accumulate, activations in tensor memory)",
                        "evals_per_s": res, "unit": "evals/s per GPU, inputs resident in HBM, wall clock around 5 steps",
                        "sign_agreement_vs_fp64_on_golden_boards": agree, "parity": "none claimed: reduced precision, see include/chessai_b200.h"}
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
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * N_BOARDS * e2e_steps / e2e_s

    # ---- training throughput (second half of BASELINE.json's metric; configs[3] shape, reported beside the headline) ---
    # data-parallel minibatch steps of Optimizer::optimize's extension: B_TRAIN synthetic boards per GPU per step,
    # teacher targets = latest.txt outputs + N(0, 0.05), NCCL all-reduce of the gradient inside the library.
    train = None
    if not args.no_train:
        dpm = importlib.import_module("chess-ai_b200.dp")
        B_TRAIN = 65536
        tcodes = dev_sets[1][:B_TRAIN]
        ttargets = torch.empty(B_TRAIN, dtype=torch.float64, device="cuda")
        net.eval_dev_ptr(tcodes.data_ptr(), B_TRAIN, ttargets.data_ptr(), main_stream.cuda_stream)
        torch.cuda.synchronize()
        gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
        ttargets = (ttargets + 0.05 * torch.randn(B_TRAIN, dtype=torch.float64, device="cuda", generator=gen)).clamp_(-4, 4)
        student = cab.Network.randomNetwork(DEFAULT_DIMS, device=local_rank, seed=1234)   # same seed: identical replicas
        student.set_weights(student.get_weights() * 0.1)
        par = dpm.DataParallel(cab, student) if world > 1 else None
        lam, accepts, t_steps = 0.5, 0, 8

        def train_step(lam):
            if par is not None:
                return par.train_batch_dev(tcodes.data_ptr(), ttargets.data_ptr(), B_TRAIN, lam)
            import ctypes as C
            lam_c, acc, e0, e1 = C.c_double(lam), C.c_int(0), C.c_double(0), C.c_double(0)
            cab._ck(cab.lib().cab_train_batch_dev(student._h, tcodes.data_ptr(), ttargets.data_ptr(), B_TRAIN, C.byref(lam_c),
                                                  15.0, 1.1, C.byref(acc), C.byref(e0), C.byref(e1), None))
            return lam_c.value, bool(acc.value), e0.value, e1.value

        for _ in range(2):
            lam, acc, e_first, _ = train_step(lam)
        l0 = student.launch_count
        barrier()
        t0 = time.perf_counter()
        for _ in range(t_steps):
            lam, acc, e0, e1 = train_step(lam)
            accepts += int(acc)
            if lam < 1e-4 or lam > 10.0:
                lam = 2.0
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        train = {"value": world * B_TRAIN * t_steps / dt, "unit": "samples/s", "batch_per_gpu": B_TRAIN, "steps": t_steps,
                 "ms_per_step": dt / t_steps * 1e3, "accepted_steps": accepts, "E_first": e_first, "E_last": e1 if acc else e0,
                 "flop_per_sample": 577944, "gpu_launches": int(student.launch_count - l0),
                 "collective": "ncclAllReduce(64220 fp64) + 1 scalar per step" if world > 1 else "none"}
        if par is not None:
            par.shutdown()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (mlp_forward_kernel) on the fp64 tensor pipe -------------------------------
    peak = cab.measure_pipe_peak("dmma", local_rank)
    n_fwd_launches = args.steps * (N_BOARDS // BATCH)
    avg_launch_s = (ms * 1e-3) / n_fwd_launches          # launches overlap on %d streams: region time / launches
    achieved = FLOP_PER_EVAL * BATCH / avg_launch_s * 1e-12
    traffic = traffic_steady = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("mlp_forward_kernel_dram_bytes_per_launch")
            traffic_steady = tj.get("mlp_forward_kernel_dram_bytes_per_launch_steady_state")
        except Exception:
            traffic = None
    roofline = {"bound": "tensor", "pipe": "fp64 tensor pipe (DMMA.8x8x4)", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "traffic_steady_state": traffic_steady,
                "algorithmic_bytes_per_launch": 72 * BATCH,
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
                        "sample": "%d of the same synthetic boards x 4 passes, one Network::clone() per thread" % CPU_SAMPLE_BOARDS,
                        "max_rel_diff_vs_gpu": agree}
        try:  # configs[0] also asks for the single-thread rate
            r1, _, _, _ = cpu_reference_rate(codes[:4096], 1, 1)
            cpu_baseline["value_1_thread"] = r1
        except Exception as exc:
            cpu_baseline["value_1_thread_error"] = str(exc)
        try:  # Optimizer::optimize is inherently sequential per network: 1 thread
            from oracle_bind import RefOracle
            if RefOracle.available():
                rnet = RefOracle().net_from_weights(DEFAULT_DIMS, latest_weights())
                n_t = 4096
                secs, _ = rnet.time_train(codes[:n_t], np.zeros(n_t), 2.0)
                cpu_baseline["train_value"] = n_t / secs
                cpu_baseline["train_unit"] = "samples/s"
                cpu_baseline["train_sample"] = "%d Optimizer::optimize steps, 1 thread" % n_t
        except Exception as exc:  # the eval baseline stands on its own
            cpu_baseline["train_error"] = str(exc)

    line = {
        "metric": "MLP board evals/sec", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(world),
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": N_BOARDS * 64, "d2h_bytes_per_step": N_BOARDS * 8,
                "steps": e2e_steps, "api": "cab_eval_host (pinned host buffers)"},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": sampler.summary(),
        "parity_max_rel_err_vs_reference_golden": parity, "streams": len(streams), "train": train,
        "precision_paths": {"f16_tcgen05": f16_path},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
