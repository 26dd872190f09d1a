"""Records games with the reference's OWN rules engine (oracle/_ref, ref_record_game in oracle/ref_harness.cpp) into
tests/golden/chess_games.npz: per ply the board codes, side to move, no-capture counter, the full legal move list in the
reference's generation order and the move played; per game the result. Run where /root/reference exists:
    python tests/golden/make_chess_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libchessai_ref.so"))
i8p = np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
L.ref_record_game.restype = C.c_int
L.ref_record_game.argtypes = [C.c_char_p, C.c_uint, C.c_int, i8p, i8p, i32p, i32p, C.c_int, i32p, i32p,
                              C.POINTER(C.c_int), C.POINTER(C.c_int)]
L.ref_perft.restype = C.c_long
START = os.path.join(HERE, "start_position.txt").encode()
N_GAMES, MAX_PLIES, CAP = int(sys.argv[1]) if len(sys.argv) > 1 else 48, 160, 160 * 80

codes, side, nocap, nmoves, moves, chosen, game_of, results, overs, plies = [], [], [], [], [], [], [], [], [], []
for g in range(N_GAMES):
    c = np.zeros((MAX_PLIES + 1) * 64, dtype=np.int8)
    s = np.zeros(MAX_PLIES + 1, dtype=np.int8)
    nm = np.zeros(MAX_PLIES + 1, dtype=np.int32)
    mv = np.zeros(CAP * 5, dtype=np.int32)
    ch = np.zeros(MAX_PLIES + 1, dtype=np.int32)
    nc = np.zeros(MAX_PLIES + 1, dtype=np.int32)
    res, over = C.c_int(0), C.c_int(0)
    n = L.ref_record_game(START, 1000 + g, MAX_PLIES, c, s, nm, mv, CAP, ch, nc, C.byref(res), C.byref(over))
    codes.append(c.reshape(-1, 64)[:n + 1]); side.append(s[:n + 1]); nocap.append(nc[:n + 1]); nmoves.append(nm[:n + 1])
    total = int(nm[:n].sum())
    moves.append(mv.reshape(-1, 5)[:total]); chosen.append(ch[:n]); game_of += [g] * (n + 1)
    results.append(res.value); overs.append(over.value); plies.append(n)
    print("game", g, "plies", n, "over", over.value, "result", res.value, flush=True)
mv_all = np.concatenate(moves)
special = {"promotions": int((mv_all[:, 4] != 0).sum())}
np.savez_compressed(os.path.join(HERE, "chess_games.npz"), codes=np.concatenate(codes), side=np.concatenate(side),
                    nocap=np.concatenate(nocap), nmoves=np.concatenate(nmoves), moves=mv_all.astype(np.int8),
                    chosen=np.concatenate(chosen), game_of=np.array(game_of, dtype=np.int32), results=np.array(results, dtype=np.int8),
                    overs=np.array(overs, dtype=np.int8), plies=np.array(plies, dtype=np.int32),
                    perft=np.array([L.ref_perft(START, d) for d in (1, 2, 3)], dtype=np.int64))
print("positions", sum(plies) + N_GAMES, "moves", len(mv_all), special, os.path.getsize(os.path.join(HERE, "chess_games.npz")), "bytes")
