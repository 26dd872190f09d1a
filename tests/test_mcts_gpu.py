"""GPU tests of the device-resident MCTS self-play (csrc/cab_mcts.cu, BASELINE.json configs[2]).

  * the kernels' move generator (csrc/chess_rules.cuh spread over a warp) against every position of the games recorded
    from the reference's own rules code: identical legal-move lists in identical order — integer work, bit-exact;
    and the "move leaves the opponent no reply" flag (decider.cpp:47-52) against the host rules engine
  * the search against tests/golden/mcts_visits.npz, recorded from the reference's OWN tree::MCTS::mcts() (Node +
    expand_node + mcts, tree.cpp:202-306, latest.txt as Decider, no root noise): with one playout in flight and one
    root the device search must walk the SAME tree — visit counts identical, priors and value sums equal
  * with many playouts in flight (virtual loss) and several roots: the tree invariants
  * whole games: legal, finished, training cases with targets inside the rule's range
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import DEFAULT_DIMS, GOLDEN, ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def net(cab, golden):
    return cab.Network.from_weights(DEFAULT_DIMS, golden["latest_weights"])


def test_device_move_generator_matches_reference_games(cab):
    g = np.load(os.path.join(GOLDEN, "chess_games.npz"))
    codes, side, nmoves, moves = g["codes"], g["side"], g["nmoves"], g["moves"].astype(np.int32)
    counts, got = cab.device_movegen(codes, side, queen_only=False)
    assert np.array_equal(counts, nmoves)          # every position, the final ones of the games included
    at = p = 0
    for n_plies in g["plies"]:                     # the move LISTS are recorded for the plies that were played
        for _ in range(int(n_plies)):
            n = int(nmoves[p])
            want = moves[at:at + n]
            m = got[p, :n].astype(np.int32)
            assert np.array_equal(m[:, 0], want[:, 0] * 8 + want[:, 1]) and np.array_equal(m[:, 1], want[:, 2] * 8 + want[:, 3]), p
            assert np.array_equal(m[:, 2], want[:, 4]), p
            at += n
            p += 1
        p += 1
    assert at == len(moves) and p == len(codes) > 7000


def test_device_no_reply_flag_matches_host_rules(cab):
    from test_fastchess import LIB
    fc = C.CDLL(LIB)
    g = np.load(os.path.join(GOLDEN, "chess_games.npz"))
    rng = np.random.default_rng(9)
    first = np.concatenate([[0], np.cumsum(g["plies"] + 1)[:-1]])
    late = [int(first[gi] + g["plies"][gi] - 1) for gi in range(len(g["plies"])) if g["overs"][gi] and g["plies"][gi] > 0]
    idx = np.array(sorted(set(late) | set(int(i) for i in rng.choice(len(g["codes"]), size=400, replace=False))))
    codes, side = np.ascontiguousarray(g["codes"][idx]), np.ascontiguousarray(g["side"][idx])
    counts, moves, flags = cab.device_movegen(codes, side, queen_only=True, no_reply=True)
    flagged = 0
    for p in range(len(codes)):
        for i in range(int(counts[p])):
            child = codes[p].copy()
            f, t, pr = (int(x) for x in moves[p, i])
            fc.cr_apply(child.ctypes.data_as(C.POINTER(C.c_byte)), f, t, 7 if pr else 0)   # the key's LAST promotion: the bishop
            want = 0 if fc.cr_has_move(child.ctypes.data_as(C.POINTER(C.c_byte)), -int(side[p])) else 1
            assert int(flags[p, i]) == want, (p, i)
            flagged += want
    assert flagged >= 5      # mates and stalemates were really among them


def test_search_equals_the_reference_mcts(cab, net):
    g = np.load(os.path.join(GOLDEN, "mcts_visits.npz"))
    n_pos, sims = len(g["codes"]), int(g["sims"])
    sp = cab.SelfPlay(net, games=n_pos, playouts_per_move=sims, roots=1, in_flight=1, max_moves=1, network_priors=0, noise=0)
    sp.set_positions(g["codes"], g["side"], g["nocap"])
    st = sp.run(search_only=True)
    assert st["searches_done"] == n_pos and st["playouts"] == n_pos * sims and st["arena_full"] == 0
    at = 0
    for p in range(n_pos):
        k = int(g["n"][p])
        rc = sp.root_children(p)
        assert len(rc["from"]) == k
        assert np.array_equal(rc["from"], g["frm"][at:at + k]) and np.array_equal(rc["to"], g["to"][at:at + k]), p
        assert rc["root_visits"] == int(g["root_visits"][p])
        assert np.array_equal(rc["visits"], g["visits"][at:at + k]), (p, rc["visits"], g["visits"][at:at + k])
        assert np.abs(rc["prior"] - g["prior"][at:at + k]).max() <= 1e-14
        assert np.abs(rc["value_sum"] - g["value_sum"][at:at + k]).max() <= 1e-9 * max(1.0, np.abs(g["value_sum"][at:at + k]).max())
        assert abs(rc["root_value_sum"] - float(g["root_value_sum"][p])) <= 1e-9 * max(1.0, abs(float(g["root_value_sum"][p])))
        at += k
    sp.close()


@pytest.mark.parametrize("network_priors", [0, 1])
def test_virtual_loss_search_keeps_the_tree_invariants(cab, net, network_priors):
    g = np.load(os.path.join(GOLDEN, "mcts_visits.npz"))
    n_pos = len(g["codes"])
    sp = cab.SelfPlay(net, games=n_pos, playouts_per_move=256, roots=4, in_flight=8, max_moves=1, network_priors=network_priors, noise=1, seed=7)
    sp.set_positions(g["codes"], g["side"], g["nocap"])
    st = sp.run(search_only=True)
    assert st["searches_done"] == n_pos and st["playouts"] == n_pos * 256 and st["arena_full"] == 0
    assert st["boards_evaluated"] > st["expansions"] >= n_pos * 4
    for p in range(n_pos):
        rc = sp.root_children(p)
        assert rc["root_visits"] == 256 and rc["visits"].sum() == 256          # every playout leaves through one child; no loss left
        assert abs(rc["prior"].sum() - 1.0) < 1e-12 and (rc["prior"] > 0).all()
        assert (rc["value_sum"] >= 0).all() and (rc["value_sum"] <= rc["visits"] + 1e-9).all()
        assert len(rc["from"]) == int(g["n"][p])
        for r in range(4):
            one = sp.root_children(p, root=r)
            assert one["root_visits"] == 64 and one["visits"].sum() == 64
    sp.close()


def test_self_play_games_are_legal_and_yield_cases(cab, net):
    from test_fastchess import LIB
    fc = C.CDLL(LIB)
    sp = cab.SelfPlay(net, games=48, playouts_per_move=48, roots=2, in_flight=4, max_moves=10, network_priors=1, noise=1, save_ratio=0.5, seed=3)
    st = sp.run()
    assert st["games_finished"] == 48 and st["moves_played"] >= 48 and st["arena_full"] == 0
    assert 0 < st["evaluator_seconds"] < st["seconds"]
    distinct = set()
    for gidx in range(48):
        s = sp.game_state(gidx)
        assert s["moves_played"] == 10 or s["over"] == 1
        assert (np.abs(s["codes"]) <= 9).all() and (s["codes"] == 1).sum() + (s["codes"] == 2).sum() == 1      # one white king
        assert (s["codes"] == -1).sum() + (s["codes"] == -2).sum() == 1
        assert s["side"] == (1 if s["moves_played"] % 2 == 0 else -1)
        distinct.add(s["codes"].tobytes())
    assert len(distinct) > 8                  # the root noise makes the games differ
    codes, targets = sp.cases()
    assert len(codes) == st["cases_kept"] and 100 < len(codes) < 48 * 10
    assert (np.abs(targets) <= 10.0 + 1e-12).all() and np.abs(targets).max() > 0
    sp.close()


def test_self_play_with_the_fp16_evaluator_keeps_the_search_invariants(cab, net):
    """cab_mcts_params.precision = 1: the fused fp16 tcgen05 evaluator behind the same search (no value parity claimed):
    the tree bookkeeping must still hold and the games must finish."""
    sp = cab.SelfPlay(net, games=32, playouts_per_move=64, roots=2, in_flight=8, max_moves=3, network_priors=1, noise=0, seed=5, precision=1)
    st = sp.run(search_only=True)
    assert st["searches_done"] == 32 and st["arena_full"] == 0 and st["eval_full"] == 0
    for gidx in range(32):
        rc = sp.root_children(gidx)
        assert rc["root_visits"] == 64 and rc["visits"].sum() == 64 and len(rc["from"]) == 20     # the start position's 20 moves
        assert abs(rc["prior"].sum() - 1.0) < 1e-12 and (rc["prior"] > 0).all()
    st = sp.run()
    assert st["games_finished"] == 32 and st["moves_played"] == 96
    sp.close()
    with pytest.raises(cab.CabError):
        cab.SelfPlay(net, games=1, playouts_per_move=8, precision=7)


def test_self_play_is_deterministic_for_a_seed(cab, net):
    """Trees are walked one warp per tree with the playouts of a tree in order, requests are served in whatever order the
    hardware schedules them: the GAMES must not depend on that order. Same seed -> the same positions, results and cases."""
    def play(seed):
        sp = cab.SelfPlay(net, games=24, playouts_per_move=64, roots=2, in_flight=8, max_moves=6, network_priors=1, noise=1, save_ratio=0.5, seed=seed)
        st = sp.run()
        states = [sp.game_state(i) for i in range(24)]
        codes, targets = sp.cases()
        order = np.lexsort(np.concatenate([codes.astype(np.int16), (targets[:, None] * 1e6).astype(np.int64)], axis=1).T[::-1])
        sp.close()
        return st, states, codes[order], targets[order]
    a, b, c = play(11), play(11), play(12)
    assert a[0]["playouts"] == b[0]["playouts"] and a[0]["boards_evaluated"] == b[0]["boards_evaluated"]
    for x, y in zip(a[1], b[1]):
        assert np.array_equal(x["codes"], y["codes"]) and (x["side"], x["over"], x["result"], x["moves_played"]) == (y["side"], y["over"], y["result"], y["moves_played"])
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    assert any(not np.array_equal(x["codes"], y["codes"]) for x, y in zip(a[1], c[1]))     # another seed: other games
