"""CPU tests that PIN the oracle (oracle/mlp_oracle.c) before anything trusts it:

  (a) against tests/golden/* — vectors produced by running the reference's own code (make_golden.py), and
  (b) against oracle/_ref/libchessai_ref.so (the reference's translation units) bit-for-bit, when it is built.

Integer / byte work is compared bit-exactly; the fp64 restatement follows the reference's operation order and is
compiled without FMA contraction, so it is also compared bit-exactly.
"""
import hashlib
import os

import numpy as np
import pytest

from conftest import DEFAULT_DIMS, GOLDEN


def test_code_table_is_k_times_0p4_bitwise(port, golden):
    ks = np.arange(-9, 10)
    vals = np.array([port.lib.orc_code_input(int(k)) for k in ks])
    assert np.array_equal(vals, golden["code_values"])            # reference: text parser + Piece::code()
    assert np.array_equal(vals, ks.astype(np.float64) * 0.4)       # what the CUDA kernel computes
    # as-written expression for every (type, colour, flag)
    for t in range(7):
        for c in range(3):
            for f in range(2):
                k = port.lib.orc_piece_code(t, c, f)
                assert port.lib.orc_piece_input(t, c, f) == k * 0.4


def test_case_parser_and_encoder_on_reference_fixture(port, golden):
    rec, target = port.parse_case_file(os.path.join(GOLDEN, "board_20.txt"))
    codes = port.encode_records(rec)[0]
    assert np.array_equal(codes, golden["case_codes"][20])
    assert target == golden["case_targets"][20]
    assert np.array_equal(codes.astype(np.float64) * 0.4, golden["case_inputs"][20])
    rec0, t0 = port.parse_case_file(os.path.join(GOLDEN, "start_position.txt"))
    assert t0 is None
    start = port.encode_records(rec0)[0]
    assert list(start[:8]) == [4, 6, 7, 3, 1, 7, 6, 4] and list(start[8:16]) == [8] * 8
    assert list(start[56:]) == [-4, -6, -7, -3, -1, -7, -6, -4] and np.all(start[16:48] == 0)


def test_forward_matches_reference_golden_bitwise(port, golden):
    net = port.net_from_weights(DEFAULT_DIMS, golden["latest_weights"])
    assert np.array_equal(net.predict(golden["case_codes"]), golden["case_forward"])
    assert np.array_equal(net.predict(golden["syn_codes"][:256]), golden["syn_forward_latest"][:256])
    assert golden["case_forward"][0] == 3.4593943354366274e-07     # SURVEY.md §8c pinned values
    assert golden["case_forward"][20] == 0.52498458704792217


def test_forward_other_dims_matches_reference_golden_bitwise(port, golden):
    dims = golden["json"]["random_small_dims"]
    assert dims == port.normalize_dims([40, 24])                  # Network({40,24}) -> {64,40,24,1}
    net = port.net_from_weights(dims, golden["random_small_weights"])
    assert np.array_equal(net.predict(golden["syn_codes"][:512]), golden["syn_forward_small"])


def test_dims_normalisation(port):
    assert port.normalize_dims([]) == DEFAULT_DIMS
    assert port.normalize_dims([64, 1]) == DEFAULT_DIMS
    assert port.normalize_dims([2048, 2048, 2048, 2048]) == [64, 2048, 2048, 2048, 2048, 1]
    assert port.normalize_dims([64, 10, 1]) == [64, 10, 1]


def test_mt19937_stream_and_random_network_match_reference(port, golden):
    mt = port.mt_new()
    first4 = [port.lib.orc_random(mt, 0.0, 1.0) for _ in range(4)]
    assert first4 == golden["json"]["rng"]["first4"]
    mt = port.mt_new()
    w = port.random_net(DEFAULT_DIMS, mt).get_weights()
    assert w[0] == golden["json"]["random_default"]["first_weight"]
    assert hashlib.sha256(w.tobytes()).hexdigest() == golden["json"]["random_default"]["weights_sha256"]
    w2 = port.random_net(golden["json"]["random_small_dims"], mt).get_weights()   # the stream continues
    assert np.array_equal(w2, golden["random_small_weights"])


def test_optimize_and_epoch_match_reference_golden(port, golden):
    net = port.net_from_weights(DEFAULT_DIMS, golden["latest_weights"])
    acc, lam, e0, e1 = net.optimize(golden["case_codes"][20], float(golden["case_targets"][20]), 2.0)
    g = golden["json"]["optimize_board20"]
    assert (acc, lam) == (0, g["lambda_after"])
    assert net.predict(golden["case_codes"][20:21])[0] == g["prediction_after"]
    net = port.net_from_weights(DEFAULT_DIMS, golden["latest_weights"])
    resets, lam, accepted = net.train_sequence(golden["case_codes"], golden["case_targets"], 2.0, mt=port.mt_new())
    e = golden["json"]["epoch_latest"]
    assert (resets, lam) == (e["resets"], e["final_lambda"])
    assert hashlib.sha256(net.get_weights().tobytes()).hexdigest() == e["weights_sha256"]


def test_training_trace_with_resets_matches_reference_golden(port, golden):
    # random net + 400 steps: accepts, rejects and two lambda resets whose dropout continues the SAME mt19937 stream
    mt = port.mt_new()
    net = port.random_net(DEFAULT_DIMS, mt)
    codes, targets = golden["case_codes"], golden["case_targets"]
    lam, lams, resets = 2.0, [], 0
    for s in range(400):
        r, lam, _ = net.train_sequence(codes[s % 66:s % 66 + 1], targets[s % 66:s % 66 + 1], lam, mt=mt)
        resets += r
        lams.append(lam)
    assert resets == golden["json"]["trace"]["resets"] == 2
    assert np.array_equal(np.array(lams), golden["trace_lambdas"])
    assert hashlib.sha256(net.get_weights().tobytes()).hexdigest() == golden["json"]["trace"]["weights_sha256"]


def test_dump_is_byte_exact(port, golden, tmp_path):
    net = port.net_from_weights(DEFAULT_DIMS, golden["latest_weights"])
    p = str(tmp_path / "latest.txt")
    net.dump(p)
    j = golden["json"]["latest_txt"]
    raw = open(p, "rb").read()
    assert len(raw) == j["bytes"] and hashlib.sha256(raw).hexdigest() == j["sha256"]
    assert "%016x" % port.fnv1a64_file(p) == j["fnv1a64"]
    again = port.load_net(p)
    assert again.dims == DEFAULT_DIMS and np.array_equal(again.get_weights(), golden["latest_weights"])


def test_minibatch_extension_reduces_to_optimize(port, golden):
    a = port.net_from_weights(DEFAULT_DIMS, golden["latest_weights"])
    b = port.net_from_weights(DEFAULT_DIMS, golden["latest_weights"])
    c, t = golden["case_codes"][7], float(golden["case_targets"][7])
    ra = a.optimize(c, t, 0.05)
    rb = b.train_batch(c[None], np.array([t]), 0.05)
    assert ra[:2] == rb[:2] and ra[2] == rb[2]
    assert np.abs(a.get_weights() - b.get_weights()).max() < 1e-13   # (psi*lambda)*a vs sum-then-scale rounding


def test_philox_known_answer(port):
    # Random123 known-answer test vectors for Philox4x32-10
    assert [hex(x) for x in port.philox([0, 0, 0, 0], [0, 0])] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in port.philox([0xffffffff] * 4, [0xffffffff] * 2)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(x) for x in port.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


# ---- the reference's own code, when oracle/_ref is built ------------------------------------------------------------
def test_port_equals_reference_library_bitwise(port, ref, golden):
    w = golden["latest_weights"]
    rn, pn = ref.net_from_weights(DEFAULT_DIMS, w), port.net_from_weights(DEFAULT_DIMS, w)
    codes = golden["syn_codes"][:300]
    assert np.array_equal(rn.predict(codes), pn.predict(codes))
    lam_r, lam_p = 2.0, 2.0
    for s in range(30):
        c, t = golden["case_codes"][s], float(golden["case_targets"][s])
        lam_r = rn.optimize(c, t, lam_r)
        _, lam_p, _, _ = pn.optimize(c, t, lam_p)
        assert lam_r == lam_p
    assert np.array_equal(rn.get_weights(), pn.get_weights())
    for k in range(-9, 10):
        c = np.zeros(64, dtype=np.int8)
        c[11] = k
        assert ref.codes_to_inputs(c)[11] == port.lib.orc_code_input(k)


def test_reference_library_reproduces_golden(ref, golden):
    rn = ref.net_from_weights(DEFAULT_DIMS, golden["latest_weights"])
    assert np.array_equal(rn.predict(golden["case_codes"]), golden["case_forward"])
