"""Short program for the round-2 ncu passes: a few launches of every hot kernel (forward batches of 4096, three resident
minibatch training steps of 65 536 boards, the fused fp16 evaluator, a small device-resident self-play)."""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cab = importlib.import_module("chess-ai_b200")
what = sys.argv[1:] or ["fwd", "train", "f16", "selfplay"]
w = np.fromfile(os.path.join(ROOT, "tests/golden/latest_weights.f64"), dtype="<f8")
net = cab.Network.from_weights([64, 144, 225, 100, 1], w)
codes = torch.from_numpy(cab.synthetic_boards(1 << 17)).cuda()
out = torch.empty(1 << 17, dtype=torch.float64, device="cuda")
streams = [torch.cuda.Stream() for _ in range(4)]
ptrs = [s.cuda_stream for s in streams]
if "fwd" in what:
    for _ in range(3):
        net.eval_dev_batches(codes.data_ptr(), 1 << 17, 4096, out.data_ptr(), ptrs)
    torch.cuda.synchronize()
if "train" in what:
    student = cab.Network.randomNetwork([64, 144, 225, 100, 1], seed=1234)
    student.set_weights(student.get_weights() * 0.1)
    targets = out.clone().clamp_(-4, 4)
    lam, st = cab.train_epoch_dev(student, codes.data_ptr(), targets.data_ptr(), 1 << 17, 65536, 0.5, stream=ptrs[0])
    lam, st = cab.train_epoch_dev(student, codes.data_ptr(), targets.data_ptr(), 1 << 17, 65536, lam, stream=ptrs[0])
    print("train", lam, st)
if "f16" in what:
    for _ in range(3):
        net.eval_f16_dev_batches(codes.data_ptr(), 1 << 17, 4096, out.data_ptr(), ptrs)
    torch.cuda.synchronize()
    net.eval_f16_dev_ptr(codes.data_ptr(), 1 << 17, out.data_ptr(), ptrs[0])
    torch.cuda.synchronize()
if "selfplay" in what:
    sp = cab.SelfPlay(net, games=128, playouts_per_move=200, roots=4, in_flight=16, max_moves=2, network_priors=1, noise=1, save_ratio=0.1)
    print("selfplay", sp.run())
