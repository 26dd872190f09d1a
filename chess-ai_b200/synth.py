"""Deterministic synthetic "random legal board" generator — host-side input generation only.

"Legal" = satisfies the reference's representation invariants: exactly one king per colour on non-adjacent squares,
no pawns on rows 0/7, at most 8 pawns / 2 knights / 2 bishops / 2 rooks / 1 queen per side. Flags: king/rook
`moved` ~ Bernoulli(0.5); pawn `moved2x` ~ Bernoulli(0.05) only on row 3 (white) / row 4 (black).
Output: int8 codes [n, 64], k = +-{1 K, 2 K*, 3 Q, 4 R, 5 R*, 6 N, 7 B, 8 P, 9 P*}, 0 empty, square = row*8+col.

Boards come in chunks of 65 536; chunk c draws from Philox(key = seed, counter = c), so board i does not depend on
how many boards are requested.
"""
import numpy as np

CHUNK = 65536
# slot layout of the 16 non-pawn pieces: (base code, colour sign, has moved-flag)
_SLOTS = [(1, 1, True), (1, -1, True)] + \
         [(6, 1, False)] * 2 + [(7, 1, False)] * 2 + [(4, 1, True)] * 2 + [(3, 1, False)] + \
         [(6, -1, False)] * 2 + [(7, -1, False)] * 2 + [(4, -1, True)] * 2 + [(3, -1, False)]


def _chunk(seed, c, n):
    rng = np.random.Generator(np.random.Philox(key=seed, counter=[c, 0, 0, 0]))
    codes = np.zeros((n, 64), dtype=np.int8)
    rows = np.arange(n)
    # pawns on rows 1..6
    wp = rng.integers(0, 9, n)
    bp = rng.integers(0, 9, n)
    perm48 = np.argsort(rng.random((n, 48)), axis=1) + 8
    double_step = rng.random((n, 16)) < 0.05
    for s in range(16):
        white = s < 8
        active = (s < wp) if white else (s - 8 < bp)
        sq = perm48[:, s]
        r = sq // 8
        flag = double_step[:, s] & (r == (3 if white else 4))
        val = (8 + flag.astype(np.int8)) * (1 if white else -1)
        codes[rows[active], sq[active]] = val[active].astype(np.int8)
    # the other pieces on random empty squares
    counts = {
        "wn": rng.integers(0, 3, n), "wb": rng.integers(0, 3, n), "wr": rng.integers(0, 3, n), "wq": rng.integers(0, 2, n),
        "bn": rng.integers(0, 3, n), "bb": rng.integers(0, 3, n), "br": rng.integers(0, 3, n), "bq": rng.integers(0, 2, n),
    }
    key = rng.random((n, 64))
    key[codes != 0] = np.inf
    order = np.argsort(key, axis=1)  # empty squares first, in random order
    moved = rng.random((n, 16)) < 0.5
    active_slots = np.ones((n, 16), dtype=bool)
    names = ["wn", "wn", "wb", "wb", "wr", "wr", "wq", "bn", "bn", "bb", "bb", "br", "br", "bq"]
    idx_in_kind = [0, 1, 0, 1, 0, 1, 0, 0, 1, 0, 1, 0, 1, 0]
    for s in range(2, 16):
        active_slots[:, s] = idx_in_kind[s - 2] < counts[names[s - 2]]
    squares = order[:, :16].copy()
    # kings must not touch: move the black king to the next spare empty square until they do not
    spare = 16
    for _ in range(24):
        wk, bk = squares[:, 0], squares[:, 1]
        adj = (np.abs(wk // 8 - bk // 8) <= 1) & (np.abs(wk % 8 - bk % 8) <= 1)
        if not adj.any():
            break
        squares[adj, 1] = order[adj, spare]
        spare += 1
    for s, (base, sign, has_flag) in enumerate(_SLOTS):
        act = active_slots[:, s]
        val = np.full(n, base, dtype=np.int8)
        if has_flag:
            val = val + moved[:, s].astype(np.int8)
        val = (val * sign).astype(np.int8)
        codes[rows[act], squares[act, s]] = val[act]
    return codes


def synthetic_boards(n, seed=0xC0FFEE, start_chunk=0):
    """n boards as int8 codes [n, 64]."""
    out = np.empty((n, 64), dtype=np.int8)
    done, c = 0, start_chunk
    while done < n:
        m = min(CHUNK, n - done)
        out[done:done + m] = _chunk(seed, c, CHUNK)[:m] if m < CHUNK else _chunk(seed, c, CHUNK)
        done += m
        c += 1
    return out


def codes_to_records(codes):
    """int8 codes -> 3-byte piece records (type, colour, flag), the encoder's input format."""
    t_of = np.array([0, 1, 1, 2, 3, 3, 4, 5, 6, 6], dtype=np.uint8)
    f_of = np.array([0, 0, 1, 0, 0, 1, 0, 0, 0, 1], dtype=np.uint8)
    c = np.asarray(codes, dtype=np.int8)
    a = np.abs(c.astype(np.int16))
    rec = np.zeros(c.shape + (3,), dtype=np.uint8)
    rec[..., 0] = t_of[a]
    rec[..., 1] = np.where(c == 0, 0, np.where(c > 0, 1, 2))
    rec[..., 2] = f_of[a]
    return rec
