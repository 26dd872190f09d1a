"""GPU parity of the bf16 tcgen05 path (K8, BASELINE.json configs[4]) against the fp64 oracle.

Stated tolerance (bf16 operands, fp32 accumulation, hardware tanh): |dy| <= 0.15 on the (-4, 4) output scale for
weights U(-1,1)/sqrt(fan_in); the sign of the evaluation (what move ranking uses) must agree wherever |y| >= 0.15.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ATOL = 0.15


def scaled_weights(dims, seed, gain=2.0):
    rng = np.random.default_rng(seed)
    parts = []
    for k, n in zip(dims[:-1], dims[1:]):
        parts.append((rng.uniform(-1, 1, (k, n)) * gain / np.sqrt(k)).reshape(-1))
    return np.concatenate(parts)


@pytest.mark.parametrize("dims,n", [([64, 256, 256, 1], 300), ([64, 512, 256, 1], 1000), ([64, 2048, 2048, 2048, 2048, 1], 200)])
def test_bf16_forward_vs_fp64_oracle(cab, port, golden, dims, n):
    w = scaled_weights(dims, sum(dims))
    net = cab.Network.from_weights(dims, w)
    codes = np.ascontiguousarray(np.tile(golden["syn_codes"], (2, 1))[:n])
    y = net.predict_batch_bf16(codes)
    want = port.net_from_weights(dims, w).predict(codes[:min(n, 200)])
    assert np.all(np.isfinite(y))
    err = np.abs(y[:len(want)] - want)
    assert err.max() < ATOL, err.max()
    strong = np.abs(want) >= ATOL
    assert strong.sum() > len(want) // 4
    assert np.array_equal(np.sign(y[:len(want)][strong]), np.sign(want[strong]))


def test_bf16_matches_gpu_fp64_on_a_net_both_paths_support(cab, golden):
    dims = [64, 256, 256, 1]
    w = scaled_weights(dims, 3)
    net = cab.Network.from_weights(dims, w)
    codes = golden["syn_codes"][:2048]
    y64 = net.predict_batch(codes)
    y16 = net.predict_batch_bf16(codes)
    assert np.abs(y16 - y64).max() < ATOL
    # ragged batch sizes (tile tails) and repeated calls
    for n in (1, 127, 128, 129, 257):
        assert np.abs(net.predict_batch_bf16(codes[:n]) - y64[:n]).max() < ATOL


def test_bf16_tracks_weight_updates(cab, golden):
    dims = [64, 256, 1]
    net = cab.Network.from_weights(dims, scaled_weights(dims, 5))
    codes = golden["syn_codes"][:256]
    a = net.predict_batch_bf16(codes)
    net.set_weights(scaled_weights(dims, 6))
    b = net.predict_batch_bf16(codes)
    assert np.abs(b - net.predict_batch(codes)).max() < ATOL and np.abs(a - b).max() > 0.2


def test_default_net_is_refused_by_the_bf16_path(cab, golden):
    net = cab.Network.from_weights([64, 144, 225, 100, 1], golden["latest_weights"])
    with pytest.raises(cab.CabError) as e:
        net.predict_batch_bf16(golden["syn_codes"][:8])
    assert e.value.code == cab.CAB_EINVAL


def test_wide_net_is_refused_by_the_fp64_stream_kernel(cab, golden):
    net = cab.Network([64, 2048, 2048, 1])
    with pytest.raises(cab.CabError) as e:
        net.predict_batch(golden["syn_codes"][:8])
    assert e.value.code == cab.CAB_EINVAL and "bf16" in str(e.value)


# ---- bf16 tcgen05 training step (forward with saved activations, dX / dW GEMMs) vs the fp64 rule ---------------------
# Stated tolerance: per layer, ||dW_bf16 - dW_fp64||_2 <= 3e-2 * ||dW_fp64||_2 (bf16 activations and psi, fp32
# accumulation over the batch, fp64 update on the master weights); E within 2e-2 relative.
DW_RTOL = 3e-2


def _layer_slices(dims):
    off = 0
    for k, n in zip(dims[:-1], dims[1:]):
        yield slice(off, off + k * n)
        off += k * n


@pytest.mark.parametrize("dims,n", [([64, 256, 256, 1], 300), ([64, 512, 256, 1], 1000), ([64, 256, 512, 256, 1], 64),
                                    ([64, 2048, 2048, 2048, 2048, 1], 96)])
def test_bf16_train_step_vs_fp64_oracle(cab, port, golden, dims, n):
    w = scaled_weights(dims, sum(dims) + 1)
    codes = np.ascontiguousarray(np.tile(golden["syn_codes"], (2, 1))[:n])
    rng = np.random.default_rng(n)
    targets = np.clip(rng.normal(0.0, 1.5, n), -3.9, 3.9)
    lam = 2e-3
    onet = port.net_from_weights(dims, w)
    o_acc, o_lam, o_e0, o_e1 = onet.train_batch(codes, targets, lam)
    net = cab.Network.from_weights(dims, w)
    g_lam, g_acc, g_e0, g_e1 = cab.Optimizer.train_batch_bf16(net, codes, targets, lam)
    assert abs(g_e0 - o_e0) <= 2e-2 * o_e0 and abs(g_e1 - o_e1) <= 2e-2 * o_e1
    assert g_acc == bool(o_acc) and g_lam == o_lam
    assert g_acc, "a small step must be accepted, otherwise there is no update to compare"
    dw_g, dw_o = net.get_weights() - w, onet.get_weights() - w
    for l, sl in enumerate(_layer_slices(dims)):
        num, den = np.linalg.norm(dw_g[sl] - dw_o[sl]), np.linalg.norm(dw_o[sl])
        assert den > 0 and num <= DW_RTOL * den, (l, num / den)


def test_bf16_training_reduces_the_error_and_tracks_the_fp64_trainer(cab, golden):
    dims = [64, 256, 256, 1]
    w = scaled_weights(dims, 9)
    codes = golden["syn_codes"][:2048]
    teacher = cab.Network.from_weights(dims, scaled_weights(dims, 10))
    targets = teacher.predict_batch(codes)
    a, b = cab.Network.from_weights(dims, w), cab.Network.from_weights(dims, w)
    lam_a = lam_b = 0.5
    first = last_a = last_b = None
    for _ in range(30):
        lam_a, _, e0, e1a = cab.Optimizer.train_batch_bf16(a, codes, targets, lam_a)
        lam_b, _, _, e1b = cab.Optimizer.train_batch(b, codes, targets, lam_b)
        first = e0 if first is None else first
        last_a, last_b = e1a, e1b
    assert last_a < 0.7 * first and abs(last_a - last_b) < 0.15 * last_b


def test_bf16_trainer_refuses_the_default_net(cab, golden):
    net = cab.Network.from_weights([64, 144, 225, 100, 1], golden["latest_weights"])
    with pytest.raises(cab.CabError):
        cab.Optimizer.train_batch_bf16(net, golden["syn_codes"][:64], np.zeros(64), 2.0)


def test_cta_pair_kernel_matches_single_cta(tmp_path):
    """The cta_group::2 GEMM (default from 12 288 boards up) against the single-CTA one on the same ragged problem."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = {}
    for mode in ("0", "1"):
        path = str(tmp_path / ("r%s.npz" % mode))
        env = dict(os.environ, CAB_TC_PAIR=mode)
        subprocess.run([sys.executable, os.path.join(root, "tests", "tc_pair_worker.py"), path], check=True, env=env, timeout=300)
        out[mode] = np.load(path)
    a, b = out["0"], out["1"]
    assert np.all(np.isfinite(b["y"])) and np.abs(a["y"] - b["y"]).max() < 1e-6           # same products, same k order
    assert np.array_equal(a["stats"][1:2], b["stats"][1:2]) and abs(a["stats"][2] - b["stats"][2]) < 1e-9
    dw_a, dw_b = a["w_after"] - a["w_before"], b["w_after"] - b["w_before"]
    assert np.linalg.norm(dw_b) > 0 and np.linalg.norm(dw_a - dw_b) <= 1e-5 * np.linalg.norm(dw_a)


# ---- fused fp16 tcgen05 path for {64, N1, N2, N3, 1} nets (cab_eval_f16_*, csrc/mlp_tc_fused.cuh) --------------------
# Stated tolerance (fp16 operands, fp32 accumulation, hardware tanh): |dy| <= 0.05 on the (-4, 4) scale for weights
# U(-1,1)/sqrt(fan_in) and identical sign wherever |y| >= 0.05. Not a parity path.
F16_ATOL = 0.05


@pytest.mark.parametrize("dims,n", [([64, 144, 225, 100, 1], 1000), ([64, 16, 16, 16, 1], 300), ([64, 192, 128, 256, 1], 515),
                                    ([64, 100, 37, 200, 1], 129), ([64, 144, 225, 100, 1], 1)])
def test_fused_f16_forward_vs_fp64_oracle(cab, port, golden, dims, n):
    w = scaled_weights(dims, sum(dims) + 7)
    net = cab.Network.from_weights(dims, w)
    codes = np.ascontiguousarray(np.tile(golden["syn_codes"], (2, 1))[:n])
    want = port.net_from_weights(dims, w).predict(codes)          # the CPU oracle, not the GPU's own fp64 path
    got = net.predict_batch_f16(codes)
    assert np.all(np.isfinite(got))
    assert np.abs(got - want).max() < F16_ATOL, np.abs(got - want).max()
    strong = np.abs(want) >= F16_ATOL
    assert np.array_equal(np.sign(got[strong]), np.sign(want[strong]))


def test_fused_f16_on_the_shipped_saturated_net(cab, golden):
    net = cab.Network.from_weights([64, 144, 225, 100, 1], golden["latest_weights"])
    codes = golden["syn_codes"][:2048]
    want, got = golden["syn_forward_latest"][:2048], net.predict_batch_f16(codes)   # want: recorded from the reference's own code
    strong = np.abs(want) >= 0.15
    agree = (np.sign(got[strong]) == np.sign(want[strong])).mean()
    assert strong.sum() > 1500 and agree >= 0.995, agree
    assert np.corrcoef(got, want)[0, 1] > 0.99


def test_fused_f16_tracks_weight_updates_and_refuses_other_shapes(cab, golden):
    dims = [64, 144, 225, 100, 1]
    net = cab.Network.from_weights(dims, scaled_weights(dims, 1))
    codes = golden["syn_codes"][:256]
    a = net.predict_batch_f16(codes)
    net.set_weights(scaled_weights(dims, 2))
    b = net.predict_batch_f16(codes)
    assert np.abs(b - net.predict_batch(codes)).max() < F16_ATOL and np.abs(a - b).max() > 0.2
    wide = cab.Network.from_weights([64, 512, 1], scaled_weights([64, 512, 1], 3))
    with pytest.raises(cab.CabError):
        wide.predict_batch_f16(codes)


# ---- identical first-ranked move under the reduced-precision evaluators (BASELINE.json north star) --------------------
# cab_best_move_host: all children in reduced precision, every child within tol of the reduced maximum re-evaluated in
# fp64, maximum over the refined values. With a per-board error <= tol / 2 (the stated tolerances; default tol is twice
# them) the move is PROVABLY the fp64 move, ties included. Checked on the 65 positions recorded from the reference games.
def _positions():
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "predictions.npz"))
    return g["codes"], g["side"]


def test_f16_ranking_picks_the_fp64_move(cab):
    dims = [64, 144, 225, 100, 1]
    codes, side = _positions()
    for seed in (11, 12):
        net = cab.Network.from_weights(dims, scaled_weights(dims, seed))
        refined = total = 0
        for c, s in zip(codes, side):
            want = cab.best_move(net, c, s, "fp64")
            got = cab.best_move(net, c, s, "f16")                 # default tol = 0.1 = 2 x the stated 0.05
            assert got[:2] == want[:2] and got[2] == want[2] and got[3] == want[3]
            refined += got[4]
            total += got[3]
        assert 65 <= refined < 0.5 * total                        # the rule re-evaluates a minority of the children


def test_f16_ranking_on_the_shipped_saturated_net(cab, golden):
    # latest.txt saturates most units; the fp16 path's error there is NOT bounded by the stated tolerance (measured on
    # these children: p99 0.12, max 0.75), so the proof does not apply — the recorded positions still all rank alike
    net = cab.Network.from_weights([64, 144, 225, 100, 1], golden["latest_weights"])
    codes, side = _positions()
    same = sum(cab.best_move(net, c, s, "f16")[:2] == cab.best_move(net, c, s, "fp64")[:2] for c, s in zip(codes, side))
    assert same == len(codes)


def test_bf16_ranking_picks_the_fp64_move_on_a_wide_net(cab):
    dims = [64, 512, 256, 1]
    net = cab.Network.from_weights(dims, scaled_weights(dims, 21))
    codes, side = _positions()
    refined = total = 0
    for c, s in zip(codes[:40], side[:40]):
        want = cab.best_move(net, c, s, "fp64")
        got = cab.best_move(net, c, s, "bf16")                    # default tol = 0.3 = 2 x the stated 0.15
        assert got[:2] == want[:2] and got[2] == want[2]
        refined += got[4]
        total += got[3]
    assert refined < total


def test_fused_f16_batches_replayed_as_a_graph_equal_one_launch(cab, golden):
    """cab_eval_f16_dev_batches captures its launches once per (buffers, sizes, streams) and replays the graph: the
    results equal one launch over all boards, the replay follows weight changes, other argument sets get their own graph."""
    import torch
    dims = [64, 144, 225, 100, 1]
    net = cab.Network.from_weights(dims, scaled_weights(dims, 4))
    n = 40 * 1024 + 77
    codes = torch.from_numpy(np.ascontiguousarray(np.tile(golden["syn_codes"], (21, 1))[:n])).cuda()
    out, ref = torch.zeros(n, dtype=torch.float64, device="cuda"), torch.zeros(n, dtype=torch.float64, device="cuda")
    streams = [torch.cuda.Stream() for _ in range(3)]
    ptrs = [s.cuda_stream for s in streams]
    for rep in range(3):                                   # capture, replay, replay after a weight change
        if rep == 2:
            net.set_weights(scaled_weights(dims, 5))
        net.eval_f16_dev_ptr(codes.data_ptr(), n, ref.data_ptr(), ptrs[0])
        out.zero_()
        net.eval_f16_dev_batches(codes.data_ptr(), n, 1024, out.data_ptr(), ptrs)     # 41 launches: the graph path
        torch.cuda.synchronize()
        assert torch.equal(out, ref) and float(out.abs().max()) > 0
    half = torch.zeros(n, dtype=torch.float64, device="cuda")
    net.eval_f16_dev_batches(codes.data_ptr(), 20 * 1024, 1024, half.data_ptr(), ptrs)  # another size: another graph
    torch.cuda.synchronize()
    assert torch.equal(half[:20 * 1024], ref[:20 * 1024]) and float(half[20 * 1024:].abs().max()) == 0.0
