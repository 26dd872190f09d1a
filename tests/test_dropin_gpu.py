"""Drop-in check on the GPU: the reference's OWN decider / MCTS / game code (prebuilt from /root/reference in place)
linked with integration/network_b200.cpp instead of the reference's network.cpp, driven through the reference's
public API (NetworkStorage::initialize, Decider::prediction, load_training_case + Optimizer::optimize, operator<<,
MCTS::run_mcts_multithreaded). Expected values are the reference CPU network's (tests/golden/golden.json)."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import DEFAULT_DIMS, GOLDEN, ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "integration", "_build", "dropin_check")


@pytest.mark.skipif(not os.path.exists(BIN), reason="integration/_build/dropin_check not built (needs /root/reference)")
def test_reference_engine_runs_on_the_b200_network(cab, golden, tmp_path):
    latest = str(tmp_path / "latest.txt")
    cab.Network.from_weights(DEFAULT_DIMS, golden["latest_weights"]).dump(latest)
    dump = str(tmp_path / "after.txt")
    out = subprocess.check_output([BIN, latest, os.path.join(GOLDEN, "start_position.txt"),
                                   os.path.join(GOLDEN, "board_20.txt"), dump, "12"], cwd=str(tmp_path), timeout=300)
    r = json.loads(out.decode().strip().splitlines()[-1])
    g = golden["json"]
    want = g["start_prediction"]
    assert abs(r["eval"] - want["eval"]) <= 1e-5 * abs(want["eval"])          # 3.459e-07, 1e-5 relative
    assert r["n_moves"] == want["n_moves"] == 20
    assert r["min_logit"] == r["max_logit"] == 20.0                           # decider.cpp:52-57
    assert r["lambda_after"] == g["optimize_board20"]["lambda_after"]         # step rejected: 2.0 / 1.1
    assert abs(r["eval_after"] - want["eval"]) <= 1e-5 * abs(want["eval"])    # rollback restored the network
    assert r["mcts_root_children"] == 20                                      # the reference's MCTS searched on it
    assert os.path.getsize(dump) == g["latest_txt"]["bytes"]
    again = cab.Network.loadFromFile(dump)
    assert np.abs(again.get_weights() - golden["latest_weights"]).max() < 1e-12


@pytest.mark.skipif(not os.path.exists(BIN), reason="integration/_build/dropin_check not built (needs /root/reference)")
def test_checkpoint_rotation_through_the_shim(cab, golden, tmp_path):
    """NetworkStorage with SAVE_NETWORKS = true, driven through the reference's own header (network.cpp:345-364):
    initialize(path) writes network_dump/latest.txt, saveNetwork() renames it to gen-0.txt and writes the current network,
    flushStorage() rotates once more. Every file is the reference's text format, byte for byte."""
    latest = str(tmp_path / "in.txt")
    cab.Network.from_weights(DEFAULT_DIMS, golden["latest_weights"]).dump(latest)
    os.mkdir(str(tmp_path / "network_dump"))
    subprocess.check_output([BIN, latest, os.path.join(GOLDEN, "start_position.txt"), os.path.join(GOLDEN, "board_20.txt"),
                             str(tmp_path / "after.txt"), "0", "rotate"], cwd=str(tmp_path), timeout=300)
    d = tmp_path / "network_dump"
    assert sorted(os.listdir(str(d))) == ["gen-0.txt", "gen-1.txt", "latest.txt"]
    assert (d / "gen-0.txt").read_bytes() == open(latest, "rb").read()          # what initialize(path) saved: the loaded network
    assert (d / "gen-1.txt").read_bytes() == (d / "latest.txt").read_bytes()    # saveNetwork(), then flushStorage(): same network
    assert (d / "latest.txt").read_bytes() == (tmp_path / "after.txt").read_bytes()   # operator<< of the clone
    w = cab.Network.loadFromFile(str(d / "latest.txt")).get_weights()
    assert np.abs(w - golden["latest_weights"]).max() < 1e-12                   # the rejected step's W += d; W -= d rounding only
