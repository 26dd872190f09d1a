"""GPU tests of the MCTS leaf queue with virtual-loss batching and K7, the prior softmax.

  * queue: boards pushed from many host threads come back, by slot, equal to the batched evaluator's results
  * K7: exp(cm*logit)/sum per node vs the oracle's restatement of tree.cpp:299-306
  * integration/_build/batched_mcts (prebuilt from the reference's own rules code): with one playout in flight the
    batched searcher reproduces the reference's sequential tree::MCTS::mcts() visit counts, value sums and priors
    EXACTLY; with many in flight it batches thousands of boards per flush and keeps the tree invariants.
"""
import ctypes as C
import json
import os
import subprocess
import threading

import numpy as np
import pytest

from conftest import DEFAULT_DIMS, GOLDEN, ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "integration", "_build", "batched_mcts")


def test_leaf_queue_near_capacity_pushes_never_share_a_slot(cab, golden):
    # many threads racing for the last slots: every successful reservation is disjoint, failures leave no hole,
    # pending never exceeds the capacity (the reservation is a compare-and-swap, not fetch_add + fetch_sub)
    net = cab.Network.from_weights(DEFAULT_DIMS, golden["latest_weights"])
    lib = cab.lib()
    cap = 1000
    for _ in range(20):
        q = C.c_void_p(None)
        cab._ck(lib.cab_leafq_create(net._h, cap, 4096, 2, C.byref(q)))
        first = C.c_long(-1)
        filler = np.zeros((cap - 37, 64), dtype=np.int8)
        cab._ck(lib.cab_leafq_push(q, filler.reshape(-1), cap - 37, C.byref(first)))
        got, over = [], []

        def worker(t):
            rng = np.random.default_rng(t)
            for _ in range(40):
                n = int(rng.integers(1, 12))
                slot = C.c_long(-1)
                rc = lib.cab_leafq_push(q, np.full((n, 64), t + 1, dtype=np.int8).reshape(-1), n, C.byref(slot))
                if rc == 0:
                    got.append((slot.value, n))
                else:
                    assert rc == cab.CAB_ESTATE
                pend = C.c_long(0)
                lib.cab_leafq_pending(q, C.byref(pend))
                over.append(pend.value)

        ts = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert max(over) <= cap
        got.sort()
        at = cap - 37
        for slot, n in got:                      # contiguous, disjoint, in order of reservation
            assert slot == at
            at += n
        assert at <= cap
        pend = C.c_long(0)
        cab._ck(lib.cab_leafq_pending(q, C.byref(pend)))
        assert pend.value == at
        cab._ck(lib.cab_leafq_destroy(q))


def test_leaf_queue_from_threads_matches_batched_eval(cab, golden):
    net = cab.Network.from_weights(DEFAULT_DIMS, golden["latest_weights"])
    lib = cab.lib()
    q = C.c_void_p(None)
    cab._ck(lib.cab_leafq_create(net._h, 20000, 4096, 4, C.byref(q)))
    codes = golden["syn_codes"]
    want = net.predict_batch(codes)
    tickets = {}

    def worker(t):
        for k in range(t, 256, 8):                      # 8 boards per push, interleaved over threads
            slot = C.c_long(-1)
            chunk = np.ascontiguousarray(codes[k * 8:(k + 1) * 8])
            cab._ck(lib.cab_leafq_push(q, chunk.reshape(-1), 8, C.byref(slot)))
            tickets[k] = slot.value

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    pend = C.c_long(0)
    cab._ck(lib.cab_leafq_pending(q, C.byref(pend)))
    assert pend.value == 2048 and sorted(tickets.values()) == list(range(0, 2048, 8))
    n = C.c_long(0)
    cab._ck(lib.cab_leafq_flush(q, C.byref(n)))
    assert n.value == 2048
    res = C.POINTER(C.c_double)()
    cab._ck(lib.cab_leafq_results(q, C.byref(res)))
    got = np.ctypeslib.as_array(res, shape=(2048,))
    for k, slot in tickets.items():
        assert np.array_equal(got[slot:slot + 8], want[k * 8:(k + 1) * 8])
    cab._ck(lib.cab_leafq_pending(q, C.byref(pend)))
    assert pend.value == 0
    # overflow is reported, not silently dropped
    big = np.zeros((20001, 64), dtype=np.int8)
    slot = C.c_long(0)
    assert lib.cab_leafq_push(q, big.reshape(-1), 20001, C.byref(slot)) == cab.CAB_ESTATE
    cab._ck(lib.cab_leafq_destroy(q))


def test_prior_softmax_vs_oracle(cab, port):
    rng = np.random.default_rng(0)
    sizes = [1, 20, 37, 3, 64, 218]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    logits = rng.uniform(-4, 4, off[-1])
    logits[off[1]:off[2]] = 20.0                           # the reference's constant-logit case: uniform priors
    cm = np.array([1.0, -1.0, 1.0, -1.0, 1.0, -1.0])
    out = np.zeros(off[-1])
    cab._ck(cab.lib().cab_prior_softmax_host(0, logits, off, len(sizes), cm, out))
    for s in range(len(sizes)):
        want = port.prior_softmax(logits[off[s]:off[s + 1]], cm[s])
        got = out[off[s]:off[s + 1]]
        assert np.abs(got - want).max() <= 1e-14
        assert abs(got.sum() - 1.0) < 1e-12
    assert np.allclose(out[off[1]:off[2]], 1.0 / 20)


def run(cab, golden, tmp_path, *args):
    latest = str(tmp_path / "latest.txt")
    if not os.path.exists(latest):
        cab.Network.from_weights(DEFAULT_DIMS, golden["latest_weights"]).dump(latest)
    out = subprocess.check_output([BIN, latest, os.path.join(GOLDEN, "start_position.txt")] + [str(a) for a in args],
                                  cwd=str(tmp_path), timeout=600)
    return json.loads(out.decode().strip().splitlines()[-1])


@pytest.mark.skipif(not os.path.exists(BIN), reason="integration/_build/batched_mcts not built (needs /root/reference)")
def test_one_playout_in_flight_equals_the_reference_search(cab, golden, tmp_path):
    r = run(cab, golden, tmp_path, 1, 40, 1, "parity")
    assert r["invariants_ok"] == 1 and r["parity_checked"] == 1
    assert r["parity_equal"] == 1


@pytest.mark.skipif(not os.path.exists(BIN), reason="integration/_build/batched_mcts not built (needs /root/reference)")
def test_virtual_loss_batches_many_boards_per_flush(cab, golden, tmp_path):
    r = run(cab, golden, tmp_path, 16, 24, 64, "bench", 1)      # network-driven priors: parent + every child board
    assert r["invariants_ok"] == 1
    assert r["boards_evaluated"] > 5000
    assert r["avg_boards_per_flush"] > 300 and r["max_boards_per_flush"] >= 16 * 21
    assert r["flushes"] < r["expansions"] / 8                   # many expansions share one GPU round trip


def test_selfplay_driver(cab, golden, tmp_path):
    """chess-ai_b200/host/selfplay.cpp, the command-line driver of the device-resident search: bookkeeping must add up."""
    exe = os.path.join(ROOT, "chess-ai_b200", "lib", "selfplay")
    if not os.path.exists(exe):
        pytest.skip("selfplay not built")
    latest = str(tmp_path / "latest.txt")
    cab.Network.from_weights(DEFAULT_DIMS, golden["latest_weights"]).dump(latest)
    out = subprocess.check_output([exe, latest, "32", "60", "2", "-", "8", "1"], timeout=300)
    r = json.loads(out.decode().strip().splitlines()[-1])
    assert r["moves_played"] == 32 * 2 and r["playouts"] == 32 * 2 * 60 and r["games_finished"] == 32
    assert r["boards_evaluated"] > r["expansions"] * 15           # parent + ~20-30 children per expansion
    assert r["avg_boards_per_flush"] > 500 and r["flushes"] < r["expansions"] / 10
    assert r["arena_full"] == 0 and r["eval_full"] == 0
    # the reference's default search shape: 4 root-parallel trees of 125 playouts, merged like Node::combineNodes
    out4 = subprocess.check_output([exe, latest, "16", "500", "2", "-", "8", "1", "0", "7", "-", "0", "fp64", "4"], timeout=300)
    r4 = json.loads(out4.decode().strip().splitlines()[-1])
    assert r4["roots"] == 4 and r4["moves_played"] == 16 * 2 and r4["playouts"] == 16 * 2 * 4 * 125
    out0 = subprocess.check_output([exe, latest, "8", "40", "1", "-", "4", "0"], timeout=300)   # shipped inverted priors
    r0 = json.loads(out0.decode().strip().splitlines()[-1])
    assert r0["playouts"] == 8 * 40 and r0["boards_evaluated"] == r0["expansions"]               # only the parent board


def test_leaf_queue_reduced_precision_option(cab, golden):
    """cab_leafq_set_precision(1): the queue evaluates with the fused fp16 tcgen05 kernel (no parity claim: tolerance)."""
    import ctypes as C
    L = cab.lib()
    dims = [64, 144, 225, 100, 1]
    rng = np.random.default_rng(4)
    w = np.concatenate([(rng.uniform(-1, 1, (k, n)) * 2.0 / np.sqrt(k)).reshape(-1) for k, n in zip(dims[:-1], dims[1:])])
    net = cab.Network.from_weights(dims, w)
    codes = np.ascontiguousarray(golden["syn_codes"][:1500])
    q = C.c_void_p()
    cab._ck(L.cab_leafq_create(net._h, 4096, 512, 2, C.byref(q)))
    cab._ck(L.cab_leafq_set_precision(q, 1))
    slot, n_eval, res = C.c_long(-1), C.c_long(0), C.POINTER(C.c_double)()
    cab._ck(L.cab_leafq_push(q, codes.reshape(-1), len(codes), C.byref(slot)))
    cab._ck(L.cab_leafq_flush(q, C.byref(n_eval)))
    cab._ck(L.cab_leafq_results(q, C.byref(res)))
    got = np.ctypeslib.as_array(res, shape=(len(codes),)).copy()
    want = net.predict_batch(codes)
    assert n_eval.value == len(codes) and slot.value == 0
    assert np.abs(got - want).max() < 0.05
    with pytest.raises(cab.CabError):
        cab._ck(L.cab_leafq_set_precision(q, 7))
    cab._ck(L.cab_leafq_destroy(q))


def test_leaf_queue_async_flush_and_wait(cab, golden):
    """cab_leafq_flush_async + cab_leafq_wait = cab_leafq_flush; a second async flush before the wait is refused."""
    import ctypes as C
    L = cab.lib()
    net = cab.Network.from_weights([64, 144, 225, 100, 1], golden["latest_weights"])
    codes = np.ascontiguousarray(golden["syn_codes"][:5000])
    q = C.c_void_p()
    cab._ck(L.cab_leafq_create(net._h, 8192, 4096, 2, C.byref(q)))
    slot, n_eval, res = C.c_long(-1), C.c_long(0), C.POINTER(C.c_double)()
    cab._ck(L.cab_leafq_push(q, codes.reshape(-1), len(codes), C.byref(slot)))
    cab._ck(L.cab_leafq_flush_async(q, C.byref(n_eval)))
    assert n_eval.value == len(codes)
    assert L.cab_leafq_flush_async(q, C.byref(n_eval)) == cab.CAB_ESTATE
    cab._ck(L.cab_leafq_wait(q))
    cab._ck(L.cab_leafq_results(q, C.byref(res)))
    got = np.ctypeslib.as_array(res, shape=(len(codes),)).copy()
    assert np.array_equal(got, net.predict_batch(codes))
    pending = C.c_long(-1)
    cab._ck(L.cab_leafq_pending(q, C.byref(pending)))
    assert pending.value == 0
    cab._ck(L.cab_leafq_wait(q))                      # nothing in flight: a no-op
    cab._ck(L.cab_leafq_destroy(q))
