import importlib, os, sys, ctypes as C
import numpy as np
sys.path.insert(0, '.')
cab = importlib.import_module("chess-ai_b200")
g = np.load('tests/golden/predictions.npz')
w = np.fromfile('tests/golden/latest_weights.f64')
net = cab.Network.from_weights([64,144,225,100,1], w)
F = C.CDLL('chess-ai_b200/lib/libfastchess.so')
buf = (C.c_ubyte * 768)()
boards = []
for p in range(len(g['codes'])):
    b = np.ascontiguousarray(g['codes'][p])
    n = F.cr_moves(b.ctypes.data_as(C.POINTER(C.c_byte)), int(g['side'][p]), buf, 1)
    for i in range(n):
        c = b.copy(); F.cr_apply(c.ctypes.data_as(C.POINTER(C.c_byte)), buf[3*i], buf[3*i+1], 7 if buf[3*i+2] else 0); boards.append(c)
boards = np.array(boards, dtype=np.int8)
y64 = net.predict_batch(boards); y16 = net.predict_batch_f16(boards)
err = np.abs(y64 - y16)
print("children boards", len(boards), "f16 abs err: max %.3g p99 %.3g p90 %.3g median %.3g" % (err.max(), np.quantile(err,.99), np.quantile(err,.9), np.median(err)))
syn = cab.synthetic_boards(65536)
e2 = np.abs(net.predict_batch(syn) - net.predict_batch_f16(syn))
print("synthetic f16 abs err: max %.3g p999 %.3g p99 %.3g median %.3g" % (e2.max(), np.quantile(e2,.999), np.quantile(e2,.99), np.median(e2)))
for tol in (0.1, 0.5, 1.0, 2.0, 4.0):
    same = ref = 0; refined = 0; total = 0
    for p in range(len(g['codes'])):
        a = cab.best_move(net, g['codes'][p], g['side'][p], "fp64")
        b = cab.best_move(net, g['codes'][p], g['side'][p], "f16", tol)
        same += a[:2] == b[:2]; refined += b[4]; total += b[3]
    print("tol", tol, "same argmax", same, "/", len(g['codes']), "refined", refined, "of", total)
