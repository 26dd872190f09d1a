"""The N > 1 path: world_size-2 gloo test of the host-side logic on CPU, and a 2-GPU NCCL test (skipped with < 2 GPUs)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def run_workers(mode, out, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dp_worker.py"), "--mode", mode, "--out", out]
    subprocess.check_call(cmd, cwd=ROOT, timeout=600)
    return [json.load(open("%s.%d.json" % (out, r))) for r in range(2)]


def test_shard_range_covers_everything(cab):
    from importlib import import_module
    dp = import_module("chess-ai_b200.dp")
    for n in (0, 1, 7, 8, 1001, 1 << 20):
        for world in (1, 2, 3, 8):
            spans = [dp.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_gloo_world2_host_logic(tmp_path):
    r = run_workers("gloo", str(tmp_path / "gloo"), 29541)
    assert r[0]["spans"] == r[1]["spans"] == [[0, 501], [501, 1001]]
    assert r[0]["ident_ok"] and r[1]["ident_ok"]
    assert r[0]["mean_err"] == r[1]["mean_err"]                       # identical decision input on every rank
    assert abs(r[0]["mean_err"] - r[0]["whole_err"]) <= 1e-12 * r[0]["whole_err"]


@pytest.mark.gpu
def test_nccl_world2_training_matches_single_gpu(cab, tmp_path):
    if cab.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = run_workers("nccl", str(tmp_path / "nccl"), 29542)
    assert r[0]["w_sha"] == r[1]["w_sha"]                             # replicas stay bit-identical
    assert r[0]["trace"] == r[1]["trace"]
    for (lam, acc, e0, e1), (lam_s, acc_s, e0_s, e1_s) in zip(r[0]["trace"], r[0]["trace_single"]):
        assert (lam, acc) == (lam_s, acc_s)
        assert abs(e0 - e0_s) <= 1e-10 * abs(e0_s) and abs(e1 - e1_s) <= 1e-8 * abs(e1_s)
    assert r[0]["max_abs_w_diff_vs_single"] < 1e-10
