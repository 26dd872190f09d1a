"""A/B of the fused fp16 kernel: run with and without CAB_F16_SERIAL=1; prints max |difference| against the fp64 path on
the golden boards and the throughput at a few launch sizes."""
import importlib, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cab = importlib.import_module("chess-ai_b200")
w = np.fromfile(os.path.join(ROOT, "tests/golden/latest_weights.f64"), dtype="<f8")
g = np.load(os.path.join(ROOT, "tests/golden/golden.npz"))
res = {"streams": int(os.environ.get("F16_STREAMS", "4"))}
for name, net in (("latest", cab.Network.from_weights([64, 144, 225, 100, 1], w)),
                  ("small", None)):
    if net is None:
        net = cab.Network.randomNetwork([64, 144, 225, 100, 1], seed=7)
        ww = net.get_weights(); net.set_weights(ww * 0.08)
    codes = cab.synthetic_boards(5000)
    for n in (1, 127, 128, 129, 1000, 5000):
        y16 = net.predict_batch_f16(codes[:n]); y64 = net.predict_batch(codes[:n])
        res["%s_maxdiff_%d" % (name, n)] = float(np.abs(y16 - y64).max())
net = cab.Network.from_weights([64, 144, 225, 100, 1], w)
N = 1 << 21
codes = torch.from_numpy(np.tile(cab.synthetic_boards(1 << 19), (4, 1))).cuda()
out = torch.empty(N, dtype=torch.float64, device="cuda")
streams = [torch.cuda.Stream() for _ in range(int(os.environ.get("F16_STREAMS", "4")))]
ptrs = [s.cuda_stream for s in streams]
ref = torch.empty(N, dtype=torch.float64, device="cuda")
net.eval_f16_dev_ptr(codes.data_ptr(), N, ref.data_ptr(), ptrs[0])
torch.cuda.synchronize()
for batch in (4096, 16384, N):
    out.zero_()
    net.eval_f16_dev_batches(codes.data_ptr(), N, batch, out.data_ptr(), ptrs)
    torch.cuda.synchronize()
    res["batches_%d_equal_one_launch" % batch] = bool(torch.equal(out, ref))
    for _ in range(2):
        net.eval_f16_dev_batches(codes.data_ptr(), N, batch, out.data_ptr(), ptrs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        net.eval_f16_dev_batches(codes.data_ptr(), N, batch, out.data_ptr(), ptrs)
    torch.cuda.synchronize()
    res["evals_per_s_batch_%d" % batch] = N * 5 / (time.perf_counter() - t0)
print(json.dumps(res))
