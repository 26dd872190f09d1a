"""The N > 1 path: world_size-2 gloo test of the host-side logic on CPU, and a 2-GPU NCCL test (skipped with < 2 GPUs)."""
import importlib
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


@pytest.mark.gpu
def test_peer_memory_world2_training_matches_single_gpu(cab, tmp_path):
    """cab_dp_peer_*: the reduction over NVLink peer memory fused into the update kernel (no NCCL in the step)."""
    if cab.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = run_workers("peer", str(tmp_path / "peer"), 29544)
    assert r[0]["w_sha"] == r[1]["w_sha"] and r[0]["trace"] == r[1]["trace"]      # identical bits and decisions on both ranks
    for (lam, acc, e0, e1), (lam_s, acc_s, e0_s, e1_s) in zip(r[0]["trace"], r[0]["trace_single"]):
        assert (lam, acc) == (lam_s, acc_s)
        assert abs(e0 - e0_s) <= 1e-10 * abs(e0_s) and abs(e1 - e1_s) <= 1e-8 * abs(e1_s)
    assert r[0]["max_abs_w_diff_vs_single"] < 1e-10
    # an epoch of resident minibatches (7 steps, the last one short), decisions on the device
    assert r[0]["w_sha_epoch"] == r[1]["w_sha_epoch"] and r[0]["epoch"] == r[1]["epoch"]
    ep, ep1 = r[0]["epoch"], r[0]["epoch_single"]
    assert ep["stats"]["steps"] == ep1["stats"]["steps"] == 7 and ep["stats"]["accepted"] == ep1["stats"]["accepted"]
    assert ep["lam"] == ep1["lam"]
    assert np.allclose(np.array(ep["trace"]), np.array(ep1["trace"]), rtol=1e-8, atol=0)
    assert r[0]["epoch_max_abs_w_diff_vs_single"] < 1e-10


@pytest.mark.gpu
def test_cpp_trainer_data_parallel_world2_matches_replicas(cab, tmp_path):
    """chess-ai_b200/lib/train in data-parallel mode: two processes, two GPUs, file rendezvous of the NCCL id;
    steps are accepted, only rank 0 reports and writes the checkpoint."""
    import json
    import subprocess
    if cab.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "chess-ai_b200", "lib", "train")
    synth = importlib.import_module("chess-ai_b200.synth")
    codes = synth.synthetic_boards(512, seed=77)
    teacher = cab.Network.from_weights([64, 144, 225, 100, 1], np.fromfile(os.path.join(root, "tests", "golden", "latest_weights.f64"), dtype="<f8"))
    targets = teacher.predict_batch(codes)
    rec = np.zeros(len(codes), dtype=np.dtype([("codes", "i1", 64), ("target", "<f8")]))
    rec["codes"], rec["target"] = codes, targets
    shard = str(tmp_path / "c.shard")
    rec.tofile(shard)
    start = str(tmp_path / "start.txt")
    student = cab.Network.randomNetwork([64, 144, 225, 100, 1], seed=21)
    student.set_weights(student.get_weights() * 0.1)
    student.dump(start)
    out = str(tmp_path / "dump")
    os.makedirs(out)
    idf = str(tmp_path / "nccl.id")
    procs = [subprocess.Popen([exe, start, shard, out, "batch", "80", "64", str(r), "5", "100", str(r), "2", idf], stdout=subprocess.PIPE)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs)
    r = json.loads(outs[0].strip().splitlines()[-1])
    assert r["world"] == 2 and r["samples"] == 80 * 64 * 2 and r["accepted"] > 0, r     # (errors are per drawn batch: not comparable)
    assert outs[1].strip() == ""                                     # only rank 0 reports and writes
    assert sorted(os.listdir(out)) == ["latest.txt", "trainer_state.txt"]
    w_dp = cab.Network.loadFromFile(os.path.join(out, "latest.txt")).get_weights()
    assert np.all(np.isfinite(w_dp)) and np.abs(w_dp - student.get_weights()).max() > 0
