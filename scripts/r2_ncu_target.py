"""Minimal launch sequence for the round-2 `ncu --set full` capture: 2 forward launches of 4096 boards, one resident training
step of 65 536 boards (8 forward+save, 8 backward, 1 gradient, 8 re-forward pieces), one fused-fp16 launch of 2^17 boards."""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cab = importlib.import_module("chess-ai_b200")
w = np.fromfile(os.path.join(ROOT, "tests/golden/latest_weights.f64"), dtype="<f8")
net = cab.Network.from_weights([64, 144, 225, 100, 1], w)
codes = torch.from_numpy(cab.synthetic_boards(1 << 17)).cuda()
out = torch.empty(1 << 17, dtype=torch.float64, device="cuda")
s = torch.cuda.Stream()
net.eval_dev_batches(codes.data_ptr(), 8192, 4096, out.data_ptr(), [s.cuda_stream])
torch.cuda.synchronize()
student = cab.Network.randomNetwork([64, 144, 225, 100, 1], seed=1234)
student.set_weights(student.get_weights() * 0.1)
targets = out.clone().clamp_(-4, 4)
print(cab.train_epoch_dev(student, codes.data_ptr(), targets.data_ptr(), 65536, 65536, 0.5, stream=s.cuda_stream))
net.eval_f16_dev_ptr(codes.data_ptr(), 1 << 17, out.data_ptr(), s.cuda_stream)
torch.cuda.synchronize()
