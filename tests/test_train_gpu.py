"""GPU parity of the training path (K2-K5) through the C ABI against the pinned oracle and the reference's golden traces.

  * cab_train_case / cab_train_sequence restate Optimizer::optimize in the reference's own operation order, so the
    accept / reject decisions and therefore lambda must match EXACTLY; weights agree to ~1e-12 (exp / reciprocal
    rounding only).
  * cab_train_batch_* is the minibatch extension; its checker is oracle/mlp_oracle.c:orc_train_batch.
Tolerances are written next to each assertion.
"""
import importlib

import numpy as np
import pytest

from conftest import DEFAULT_DIMS, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture()
def latest(cab, golden):
    return cab.Network.from_weights(DEFAULT_DIMS, golden["latest_weights"])


def test_optimize_board20_rejected_like_reference(cab, latest, golden):
    g = golden["json"]["optimize_board20"]
    lam, acc, e0, e1 = cab.Optimizer.optimize(latest, golden["case_codes"][20], float(golden["case_targets"][20]), 2.0)
    assert acc is False
    assert lam == g["lambda_after"]                      # 2.0 / 1.1, exact
    assert e1 >= e0
    y = latest.predict_batch(golden["case_codes"][20:21])[0]
    assert abs(y - g["prediction_after"]) < 1e-12        # W + dW - dW is not bit-reversible; both sides do it


def test_epoch_over_the_66_cases_matches_reference(cab, latest, golden):
    g = golden["json"]["epoch_latest"]
    lam, acc, resets = cab.Optimizer.train_sequence(latest, golden["case_codes"], golden["case_targets"], 2.0)
    assert lam == g["final_lambda"]                      # every decision identical -> identical lambda
    assert resets == g["resets"] == 0
    assert acc.sum() == 0
    w = latest.get_weights()
    assert np.abs(w[::257] - golden["epoch_weights_sample"]).max() < 1e-10
    assert abs(latest.predict_batch(golden["case_codes"][20:21])[0] - g["predict_board20"]) < 1e-10


def test_trace_with_accepts_matches_reference_until_first_reset(cab, port, golden):
    # reference: Network::randomNetwork() in a fresh process, 400 optimize steps cycling over the 66 cases
    mt = port.mt_new()
    w0 = port.random_net(DEFAULT_DIMS, mt).get_weights()
    assert np.array_equal(w0[::101], golden["random_default_weights_sample"])
    ref_l = golden["trace_lambdas"]
    net = cab.Network.from_weights(DEFAULT_DIMS, w0)
    codes, targets = golden["case_codes"], golden["case_targets"]
    lam, lams = 2.0, []
    for s in range(400):
        lam, acc, e0, e1 = cab.Optimizer.optimize(net, codes[s % 66], float(targets[s % 66]), lam)
        lams.append(lam)
        if lam < 1e-4 or lam > 10.0:   # the reference resets here and draws dropout from mt19937: streams diverge
            break
    n = len(lams)
    assert n > 20 and np.any(np.diff(lams) > 0) and np.any(np.diff(lams) < 0)   # the trace has accepts and rejects
    assert ref_l[n - 1] == 2.0                      # the reference reset lambda at exactly this step
    assert np.array_equal(np.array(lams[:n - 1]), ref_l[:n - 1])   # identical decisions up to there


def test_sequence_with_resets_matches_port_oracle_bitwise_decisions(cab, port, golden):
    # same loop, dropout = Philox(seed, reset index) on both sides -> decisions, lambda and resets identical
    mt = port.mt_new()
    pn = port.random_net(DEFAULT_DIMS, mt)
    net = cab.Network.from_weights(DEFAULT_DIMS, pn.get_weights())
    idx = np.arange(400) % 66
    codes, targets = golden["case_codes"][idx], golden["case_targets"][idx]
    resets_p, lam_p, acc_p = pn.train_sequence(codes, targets, 2.0, philox_seed=0xABCDEF)
    lam_g, acc_g, resets_g = cab.Optimizer.train_sequence(net, codes, targets, 2.0, dropout_seed=0xABCDEF)
    assert resets_p >= 1
    assert resets_g == resets_p
    assert np.array_equal(acc_g, acc_p)
    assert lam_g == lam_p
    # 400 chained steps amplify the 1-ulp exp/reciprocal differences; weights are O(1), so an absolute bound
    assert np.abs(net.get_weights() - pn.get_weights()).max() < 1e-9


def test_minibatch_of_one_equals_optimize(cab, port, golden):
    pn = port.net_from_weights(DEFAULT_DIMS, golden["latest_weights"])
    net = cab.Network.from_weights(DEFAULT_DIMS, golden["latest_weights"])
    c, t = golden["case_codes"][7:8], golden["case_targets"][7:8]
    acc_p, lam_p, e0_p, e1_p = pn.optimize(c[0], float(t[0]), 0.05)
    lam_g, acc_g, e0_g, e1_g = cab.Optimizer.train_batch(net, c, t, 0.05)
    assert (acc_g, lam_g) == (bool(acc_p), lam_p)
    assert abs(e0_g - e0_p) <= 1e-9 * abs(e0_p) and abs(e1_g - e1_p) <= 1e-9 * abs(e1_p)
    assert np.abs(net.get_weights() - pn.get_weights()).max() < 1e-9


@pytest.mark.parametrize("n,lam", [(66, 0.5), (256, 0.01), (1000, 2.0), (37, 1e-3)])
def test_minibatch_step_vs_port_oracle(cab, port, golden, n, lam):
    rng = np.random.default_rng(n)
    w = rng.uniform(-0.3, 0.3, 64216)
    pn = port.net_from_weights(DEFAULT_DIMS, w)
    net = cab.Network.from_weights(DEFAULT_DIMS, w)
    codes = golden["syn_codes"][:n]
    targets = rng.uniform(-2, 2, n)
    acc_p, lam_p, e0_p, e1_p = pn.train_batch(codes, targets, lam)
    lam_g, acc_g, e0_g, e1_g = cab.Optimizer.train_batch(net, codes, targets, lam)
    assert abs(e0_g - e0_p) <= 1e-10 * abs(e0_p)
    assert abs(e1_g - e1_p) <= 1e-8 * abs(e1_p)
    assert (acc_g, lam_g) == (bool(acc_p), lam_p)
    assert np.abs(net.get_weights() - pn.get_weights()).max() < 1e-10


def test_minibatch_descends_on_a_learnable_target(cab, golden):
    # teacher = latest.txt outputs; a random student must reduce the batch error over accepted steps
    teacher = cab.Network.from_weights(DEFAULT_DIMS, golden["latest_weights"])
    codes = golden["syn_codes"]
    targets = teacher.predict_batch(codes)
    student = cab.Network.randomNetwork(DEFAULT_DIMS, seed=7)
    student.set_weights(student.get_weights() * 0.05)
    lam, first, last, accepts = 0.5, None, None, 0
    for step in range(30):
        lam, acc, e0, e1 = cab.Optimizer.train_batch(student, codes, targets, lam)
        first = e0 if first is None else first
        if acc:
            accepts += 1
            last = e1
        if lam < 1e-4 or lam > 10:
            lam = 2.0
    assert accepts >= 5 and last < first


def test_dropout_philox_bitwise_and_rate(cab, port, golden):
    pn = port.net_from_weights(DEFAULT_DIMS, golden["latest_weights"])
    net = cab.Network.from_weights(DEFAULT_DIMS, golden["latest_weights"])
    hits_p = pn.dropout_philox(0.01, 0x5EED, 3)
    hits_g = cab.Optimizer.dropout(net, 0.01, seed=0x5EED, offset=3)
    assert hits_g == hits_p
    w = net.get_weights()
    assert np.array_equal(w, pn.get_weights())
    changed = w != golden["latest_weights"]
    assert changed.sum() == hits_g
    assert 450 < hits_g < 850                      # Binomial(64216, 0.01): mean 642, sd 25
    assert np.all(np.abs(w[changed]) < 1.0)        # replacements are U(-1, 1)
    assert abs(w[changed].mean()) < 0.1
    # a different offset touches a different set
    net2 = cab.Network.from_weights(DEFAULT_DIMS, golden["latest_weights"])
    cab.Optimizer.dropout(net2, 0.01, seed=0x5EED, offset=4)
    assert not np.array_equal(net2.get_weights() != golden["latest_weights"], changed)


def test_random_network_is_philox_uniform(cab, port):
    net = cab.Network.randomNetwork(DEFAULT_DIMS, seed=1234, stream=9)
    w = net.get_weights()
    assert w.min() >= -1.0 and w.max() < 1.0 and abs(w.mean()) < 0.02 and abs(w.std() - 3 ** -0.5) < 0.01
    for i in (0, 1, 777, 64215):
        r = port.philox([i, 0, 9, 0], [1234, 0])
        u = ((int(r[3]) << 32 | int(r[2])) >> 11) / 2.0 ** 53
        assert w[i] == 2.0 * u - 1.0


def test_full_size_batch_update_is_the_weighted_mean_of_its_halves(cab, golden):
    """Size-independent property at the bench's batch (65 536 boards, far beyond what the CPU oracle finishes quickly):
    dW is the batch MEAN of the per-sample updates, so for any split  n * dW(A u B) = n_A * dW(A) + n_B * dW(B),
    and E(A u B) is the weighted mean of E(A) and E(B). Checked to rounding (fp64 sums in different orders)."""
    synth = importlib.import_module("chess-ai_b200.synth")
    n, n_a = 65536, 40000                                   # deliberately not aligned with the 8 192-board launch pieces
    codes = synth.synthetic_boards(n, seed=123)
    teacher = cab.Network.from_weights(DEFAULT_DIMS, golden["latest_weights"])
    targets = teacher.predict_batch(codes)
    start = cab.Network.randomNetwork(DEFAULT_DIMS, seed=31)
    w0 = start.get_weights() * 0.05
    lam = 1e-4                                              # tiny step: accepted, so the update stays in the weights

    def step(sl):
        net = cab.Network.from_weights(DEFAULT_DIMS, w0)
        _, acc, e0, _ = cab.Optimizer.train_batch(net, codes[sl], targets[sl], lam)
        assert acc
        return net.get_weights() - w0, e0

    d_all, e_all = step(slice(0, n))
    d_a, e_a = step(slice(0, n_a))
    d_b, e_b = step(slice(n_a, n))
    mix = (n_a * d_a + (n - n_a) * d_b) / n
    assert np.linalg.norm(d_all) > 0 and np.linalg.norm(d_all - mix) <= 1e-9 * np.linalg.norm(d_all)
    assert abs(e_all - (n_a * e_a + (n - n_a) * e_b) / n) <= 1e-12 * e_all


def test_epoch_on_the_device_equals_step_by_step(cab, golden):
    """cab_train_epoch_dev (lambda, accept / reject, rollback and the reset + dropout rule on the device, one wait at the
    end) against the same minibatches pushed one by one through cab_train_batch_dev with the host deciding in between."""
    import ctypes as C
    import torch
    synth = importlib.import_module("chess-ai_b200.synth")
    n, batch = 6 * 4096 + 1000, 4096                         # a short last minibatch too
    codes = synth.synthetic_boards(n, seed=5)
    teacher = cab.Network.from_weights(DEFAULT_DIMS, golden["latest_weights"])
    targets = teacher.predict_batch(codes)
    w0 = cab.Network.randomNetwork(DEFAULT_DIMS, seed=77).get_weights() * 0.05
    d_codes, d_targets = torch.from_numpy(codes).cuda(), torch.from_numpy(targets).cuda()
    lam0 = 9.5                                               # close to LAMBDA_MAX: accepted steps push it over 10 -> reset + dropout

    a = cab.Network.from_weights(DEFAULT_DIMS, w0)
    lam_a, st, trace = cab.train_epoch_dev(a, d_codes.data_ptr(), d_targets.data_ptr(), n, batch, lam0, dropout_seed=99, trace=True)
    torch.cuda.synchronize()

    b = cab.Network.from_weights(DEFAULT_DIMS, w0)
    lam, accepted, resets, steps, hist = lam0, 0, 0, 0, []
    for b0 in range(0, n, batch):
        nb = min(batch, n - b0)
        lam_c, acc, e0, e1 = C.c_double(lam), C.c_int(0), C.c_double(0), C.c_double(0)
        cab._ck(cab.lib().cab_train_batch_dev(b._h, d_codes.data_ptr() + b0 * 64, d_targets.data_ptr() + b0 * 8, nb, C.byref(lam_c), 15.0, 1.1,
                                              C.byref(acc), C.byref(e0), C.byref(e1), None))
        lam = lam_c.value
        accepted += acc.value
        hist.append((e0.value, e1.value))
        if lam < 1e-4 or lam > 10.0:                         # train.cpp:54-57
            lam = 2.0
            cab.Optimizer.dropout(b, 0.01, seed=99, offset=resets)
            resets += 1
        steps += 1
    assert (st["steps"], st["accepted"], st["resets"]) == (steps, accepted, resets) and resets >= 1 and 0 < accepted
    assert lam_a == lam
    hist = np.array(hist)
    assert np.abs(trace - hist).max() <= 1e-9 * np.abs(hist).max()          # fp64 sums in atomic order: rounding only
    assert st["first_err"] == trace[0, 0] and st["last_err"] == trace[-1, 0]
    assert np.abs(a.get_weights() - b.get_weights()).max() <= 1e-9


def test_epoch_from_host_memory_equals_the_resident_epoch(cab, golden):
    """cab_train_epoch_host (what the trainer command line calls per save interval) = one upload + cab_train_epoch_dev."""
    import ctypes as C
    import torch
    synth = importlib.import_module("chess-ai_b200.synth")
    n, batch = 3 * 2048 + 300, 2048
    codes = synth.synthetic_boards(n, seed=6)
    targets = cab.Network.from_weights(DEFAULT_DIMS, golden["latest_weights"]).predict_batch(codes)
    w0 = cab.Network.randomNetwork(DEFAULT_DIMS, seed=78).get_weights() * 0.05
    a, b = cab.Network.from_weights(DEFAULT_DIMS, w0), cab.Network.from_weights(DEFAULT_DIMS, w0)
    d_codes, d_targets = torch.from_numpy(codes).cuda(), torch.from_numpy(targets).cuda()
    lam_a, st_a = cab.train_epoch_dev(a, d_codes.data_ptr(), d_targets.data_ptr(), n, batch, 1.0, dropout_seed=5)
    st_b, lam_b = cab.TrainStats(), C.c_double(1.0)
    cab._ck(cab.lib().cab_train_epoch_host(b._h, codes.reshape(-1), targets, n, batch, C.byref(lam_b), 15.0, 1.1, 1, 5, C.byref(st_b), None))
    assert lam_b.value == lam_a and st_b.as_dict()["steps"] == st_a["steps"] == 4 and st_b.as_dict()["accepted"] == st_a["accepted"]
    assert np.abs(a.get_weights() - b.get_weights()).max() <= 1e-9
