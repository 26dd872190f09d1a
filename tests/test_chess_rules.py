"""CPU tests of csrc/chess_rules.cuh — the rules the MCTS kernels run on the GPU — compiled for the host and pinned
against the games recorded from the reference's OWN rules code (tests/golden/chess_games.npz): identical legal-move
lists in identical order at every ply, identical boards after every move, the "has the side to move a reply" flag, the
check test and the material sum. Integer work, compared bit-exactly. (The same functions are checked again on the GPU
through cab_movegen_host in tests/test_mcts_gpu.py.)"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from test_fastchess import fc, new_pos, pos_fields  # noqa: F401  (fixture + helpers; builds libfastchess.so)


def _ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_byte))


def test_device_rules_replay_reference_games(fc):
    g = np.load(os.path.join(GOLDEN, "chess_games.npz"))
    codes, side, nmoves, moves, chosen = g["codes"], g["side"], g["nmoves"], g["moves"].astype(np.int32), g["chosen"]
    pos_i = move_i = ply_i = 0
    buf = (C.c_ubyte * (3 * 256))()
    positions = 0
    for game, n_plies in enumerate(g["plies"]):
        board = np.ascontiguousarray(codes[pos_i]).copy()
        for ply in range(int(n_plies)):
            assert np.array_equal(board, codes[pos_i + ply]), (game, ply)
            s = int(side[pos_i + ply])
            n = fc.cr_moves(_ptr(board), s, buf, 0)
            want = moves[move_i:move_i + int(nmoves[pos_i + ply])]
            assert n == len(want), (game, ply, n, len(want))
            got = np.frombuffer(buf, dtype=np.uint8)[:3 * n].reshape(n, 3).astype(np.int32)
            assert np.array_equal(got[:, 0], want[:, 0] * 8 + want[:, 1]) and np.array_equal(got[:, 1], want[:, 2] * 8 + want[:, 3])
            assert np.array_equal(got[:, 2], want[:, 4]), (game, ply)
            assert fc.cr_has_move(_ptr(board), s) == int(n > 0)
            nq = fc.cr_moves(_ptr(board), s, buf, 1)       # queen-only list: one entry per std::map<Move> key
            assert nq == len({(int(a), int(b)) for a, b in got[:, :2]})
            m = got[int(chosen[ply_i + ply])]
            fc.cr_apply(_ptr(board), int(m[0]), int(m[1]), int(m[2]))
            move_i += n
            positions += 1
        assert np.array_equal(board, codes[pos_i + int(n_plies)])
        final_side = int(side[pos_i + int(n_plies)])
        assert fc.cr_has_move(_ptr(board), final_side) == int(int(nmoves[pos_i + int(n_plies)]) > 0)
        pos_i += int(n_plies) + 1
        ply_i += int(n_plies)
    assert move_i == len(moves) and positions > 7000


def test_device_rules_agree_with_host_engine_on_helpers(fc):
    g = np.load(os.path.join(GOLDEN, "chess_games.npz"))
    codes, side, nocap = g["codes"], g["side"], g["nocap"]
    p = new_pos(fc)
    rng = np.random.default_rng(11)
    for i in rng.choice(len(codes), size=1500, replace=False):
        board = np.ascontiguousarray(codes[i])
        fc.fc_set(p, _ptr(board), int(side[i]), int(nocap[i]))
        for colour in (1, -1):
            assert fc.cr_material(_ptr(board), colour) == int(fc.fc_material(p, colour))
