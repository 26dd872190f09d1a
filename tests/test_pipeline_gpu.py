"""Self-play -> packed case shard -> trainer with checkpoint rotation, on the GPU through the
C-ABI: chess-ai_b200/lib/selfplay and chess-ai_b200/lib/train (host/selfplay.cpp, host/train.cpp)."""
import importlib
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "chess-ai_b200", "lib")
DIMS = [64, 144, 225, 100, 1]
REC = np.dtype([("codes", "i1", 64), ("target", "<f8")])


def run(*args):
    out = subprocess.check_output([str(a) for a in args], timeout=600)
    return json.loads(out.decode().strip().splitlines()[-1])


@pytest.fixture(scope="module")
def cab():
    return importlib.import_module("chess-ai_b200")


@pytest.fixture(scope="module")
def latest(cab, tmp_path_factory):
    w = np.fromfile(os.path.join(ROOT, "tests", "golden", "latest_weights.f64"), dtype="<f8")
    path = str(tmp_path_factory.mktemp("net") / "latest.txt")
    cab.Network.from_weights(DIMS, w).dump(path)
    return path


@pytest.fixture(scope="module")
def shard(latest, tmp_path_factory):
    path = str(tmp_path_factory.mktemp("cases") / "selfplay.shard")
    r = run(os.path.join(LIBDIR, "selfplay"), latest, 48, 40, 12, "-", 8, 1, 0, 99, path, 1.0)
    assert r["cases_written"] == r["moves_played"] == 48 * 12            # ratio 1.0 keeps every board
    return path


def test_selfplay_cases_follow_the_target_rule(shard):
    rec = np.fromfile(shard, dtype=REC)
    assert len(rec) == 48 * 12
    codes, t = rec["codes"], rec["target"]
    assert ((codes == 1) | (codes == 2)).sum(axis=1).tolist() == [1] * len(rec)      # one white king (moved or not)
    assert ((codes == -1) | (codes == -2)).sum(axis=1).tolist() == [1] * len(rec)
    assert np.abs(codes).max() <= 9 and (t >= -10.0).all() and (t <= 10.0).all()
    # target = 10*(4*b + r)/5 with b in [-1,1], r in {-1,0,1}: unfinished 12-move games have r = 0 -> |t| <= 8
    assert (np.abs(t) <= 8.0 + 1e-12).all() and t.std() > 0.0


def test_sampling_ratio(latest, tmp_path):
    path = str(tmp_path / "tenth.shard")
    r = run(os.path.join(LIBDIR, "selfplay"), latest, 64, 20, 10, "-", 4, 0, 0, 7, path)      # default ratio 0.1
    assert 20 <= r["cases_written"] <= 120 and os.path.getsize(path) == 72 * r["cases_written"]   # 640 moves * 0.1 +- 5 sigma


def test_trainer_rotates_checkpoints_and_resumes(cab, latest, shard, tmp_path):
    out = str(tmp_path / "network_dump")
    os.makedirs(out)
    r = run(os.path.join(LIBDIR, "train"), latest, shard, out, "batch", 150, 64, 0, 1, 50)
    assert not r["resumed"] and r["total_steps"] == 150 and r["samples"] == 150 * 64 and r["decisions"] == 150
    assert sorted(os.listdir(out)) == ["gen-0.txt", "gen-1.txt", "gen-2.txt", "latest.txt", "trainer_state.txt"]
    assert r["accepted"] > 0 and r["last_err"] < r["first_err"]                     # the batch error goes down
    lam, steps, resets, gen = open(os.path.join(out, "trainer_state.txt")).read().split()
    assert float(lam) == r["lambda"] and int(steps) == 150 and int(gen) == 3
    w_latest = cab.Network.loadFromFile(os.path.join(out, "latest.txt")).get_weights()
    w_gen2 = cab.Network.loadFromFile(os.path.join(out, "gen-2.txt")).get_weights()
    assert (w_latest == w_gen2).all()                                                 # flush right after the step-150 save
    r2 = run(os.path.join(LIBDIR, "train"), latest, shard, out, "batch", 50, 64, 0, 1, 50)
    assert r2["resumed"] and r2["total_steps"] == 200
    assert "gen-4.txt" in os.listdir(out) and float(open(os.path.join(out, "trainer_state.txt")).read().split()[0]) == r2["lambda"]


def test_trainer_exact_mode_is_the_reference_loop(cab, latest, tmp_path):
    """A one-case shard makes the uniform draw deterministic: the trainer must equal Optimizer.train_sequence."""
    board, target = cab.Board.load_training_case(os.path.join(ROOT, "tests", "golden", "board_20.txt"))
    codes = board.codes()
    one = str(tmp_path / "one.shard")
    np.array([(codes, target)], dtype=REC).tofile(one)
    out = str(tmp_path / "dump")
    os.makedirs(out)
    r = run(os.path.join(LIBDIR, "train"), latest, one, out, "exact", 12, 1, 0, 5, 4)
    net = cab.Network.loadFromFile(latest)
    lam, acc, resets = cab.Optimizer.train_sequence(net, np.repeat(codes[None], 12, axis=0), np.full(12, target), lam=2.0)
    assert r["lambda"] == lam and r["accepted"] == int(np.sum(acc)) and r["resets"] == resets == 0
    ref = str(tmp_path / "ref.txt")
    net.dump(ref)
    assert open(ref, "rb").read() == open(os.path.join(out, "latest.txt"), "rb").read()
    assert sorted(os.listdir(out)) == ["gen-0.txt", "gen-1.txt", "gen-2.txt", "latest.txt", "trainer_state.txt"]
