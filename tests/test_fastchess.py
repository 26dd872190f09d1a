"""CPU tests of the fast rules engine (chess-ai_b200/host/fastchess.hpp) against games recorded
from the reference's OWN rules code (tests/golden/chess_games.npz, made by tests/golden/make_chess_golden.py):
every legal-move list must be identical (same moves, same order, promotions included) and every resulting board
identical, flags included — integer work, compared bit-exactly. Plus perft 20 / 400 / 8 902."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

LIB = os.path.join(ROOT, "chess-ai_b200", "lib", "libfastchess.so")


@pytest.fixture(scope="module")
def fc():
    src = os.path.join(ROOT, "chess-ai_b200", "host", "fastchess_c.cpp")
    hdr = os.path.join(ROOT, "chess-ai_b200", "host", "fastchess.hpp")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", "-static-libstdc++",
                               "-static-libgcc", "-o", LIB, src])
    L = C.CDLL(LIB)
    L.fc_perft.restype = C.c_long
    L.fc_material.restype = C.c_double
    return L


def new_pos(fc):
    return (C.c_byte * fc.fc_sizeof_position())()


def pos_fields(buf):
    a = np.frombuffer(buf, dtype=np.int8)
    return a[:64].copy(), int(a[64]), int(a[65]), int(a[66]), int(np.frombuffer(buf, dtype=np.int32)[17])


def test_perft_matches_reference(fc):
    g = np.load(os.path.join(GOLDEN, "chess_games.npz"))
    p = new_pos(fc)
    fc.fc_start(p)
    assert [fc.fc_perft(p, d) for d in (1, 2, 3)] == [int(x) for x in g["perft"]] == [20, 400, 8902]
    assert fc.fc_perft(p, 4) == 197281


def test_start_position_matches_reference_fixture(fc, port):
    rec, _ = port.parse_case_file(os.path.join(GOLDEN, "start_position.txt"))
    p = new_pos(fc)
    fc.fc_start(p)
    assert np.array_equal(pos_fields(p)[0], port.encode_records(rec)[0])


def test_replay_of_reference_games_is_identical(fc):
    g = np.load(os.path.join(GOLDEN, "chess_games.npz"))
    codes, side, nocap, nmoves, moves, chosen = g["codes"], g["side"], g["nocap"], g["nmoves"], g["moves"].astype(np.int32), g["chosen"]
    pos_i = move_i = ply_i = 0
    seen = {"castle": 0, "ep": 0, "promo": 0, "double": 0, "mate_or_stalemate": 0}
    buf = (C.c_ubyte * (3 * 256))()
    for game, n_plies in enumerate(g["plies"]):
        p = new_pos(fc)
        fc.fc_set(p, np.ascontiguousarray(codes[pos_i]).ctypes.data_as(C.POINTER(C.c_byte)), int(side[pos_i]), int(nocap[pos_i]))
        for ply in range(int(n_plies)):
            sq, s, over, res, nc = pos_fields(p)
            assert np.array_equal(sq, codes[pos_i + ply]), (game, ply)
            assert (s, nc) == (int(side[pos_i + ply]), int(nocap[pos_i + ply])) and over == 0
            n = fc.fc_moves(p, buf, 0)
            want = moves[move_i:move_i + int(nmoves[pos_i + ply])]
            assert n == len(want), (game, ply, n, len(want))
            got = np.frombuffer(buf, dtype=np.uint8)[:3 * n].reshape(n, 3).astype(np.int32)
            assert np.array_equal(got[:, 0], want[:, 0] * 8 + want[:, 1]) and np.array_equal(got[:, 1], want[:, 2] * 8 + want[:, 3])
            assert np.array_equal(got[:, 2], want[:, 4]), (game, ply)
            m = got[int(chosen[ply_i + ply])]
            k = abs(int(sq[m[0]]))
            seen["castle"] += int(k in (1, 2) and abs(int(m[0]) % 8 - int(m[1]) % 8) == 2)
            seen["ep"] += int(k in (8, 9) and m[0] % 8 != m[1] % 8 and sq[m[1]] == 0)
            seen["promo"] += int(m[2] != 0)
            seen["double"] += int(k in (8, 9) and abs(int(m[0]) // 8 - int(m[1]) // 8) == 2)
            fc.fc_play(p, int(m[0]), int(m[1]), int(m[2]))
            move_i += n
        sq, s, over, res, nc = pos_fields(p)
        assert np.array_equal(sq, codes[pos_i + int(n_plies)])
        assert over == int(g["overs"][game]) and res == int(g["results"][game])
        seen["mate_or_stalemate"] += int(over and nc < 50)
        pos_i += int(n_plies) + 1
        ply_i += int(n_plies)
    assert move_i == len(moves)
    # the recorder biases towards the rare rules: make sure they were really exercised
    assert seen["castle"] >= 3 and seen["ep"] >= 3 and seen["promo"] >= 10 and seen["double"] >= 50, seen
    assert seen["mate_or_stalemate"] >= 1, seen


def test_material_rule(fc):
    p = new_pos(fc)
    fc.fc_start(p)
    assert fc.fc_material(p, 1) == 0.0
    sq = pos_fields(p)[0]
    sq[59] = 0                                             # remove the black queen
    fc.fc_set(p, sq.ctypes.data_as(C.POINTER(C.c_byte)), 1, 0)
    assert fc.fc_material(p, 1) == 9.0 and fc.fc_material(p, -1) == -9.0


def test_move_generation_is_colour_symmetric(fc):
    """Mirroring a position (rows reversed, colours swapped, side to move swapped) must mirror its legal moves: an
    internal consistency check of the rules engine that does not depend on the recorded games' side to move."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "chess_games.npz"))
    codes, side, nocap = g["codes"], g["side"], g["nocap"]
    rng = np.random.default_rng(3)
    p, buf = new_pos(fc), (C.c_ubyte * (3 * 256))()

    def moves_of(c, s, nc):
        fc.fc_set(p, np.ascontiguousarray(c).ctypes.data_as(C.POINTER(C.c_byte)), int(s), int(nc))
        n = fc.fc_moves(p, buf, 0)
        return {(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]) for i in range(n)}

    def flip_sq(sq):
        return (7 - (sq >> 3)) * 8 + (sq & 7)

    checked = 0
    for i in rng.choice(len(codes), size=600, replace=False):
        mirrored = (-codes[i].reshape(8, 8)[::-1]).reshape(64)
        a = moves_of(codes[i], side[i], nocap[i])
        b = moves_of(mirrored, -side[i], nocap[i])
        assert {(flip_sq(f), flip_sq(t), pr) for f, t, pr in a} == b, int(i)
        checked += len(a)
    assert checked > 10000
