"""Times resident training epochs (cab_train_epoch_dev, 65 536 boards per step) with CUDA events: the A/B harness for the
gradient kernel's item order / slice size (CAB_GRAD_SLICE, CAB_GRAD_TASK_MAJOR are read once per process).
usage: python scripts/grad_ab.py [steps] [batch]"""
import importlib, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cab = importlib.import_module("chess-ai_b200")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 32
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
n = steps * B
teacher = cab.Network.randomNetwork([64, 144, 225, 100, 1], seed=7)
student = cab.Network.randomNetwork([64, 144, 225, 100, 1], seed=3)
student.set_weights(student.get_weights() * 0.1)
codes = torch.from_numpy(cab.synthetic_boards(4 * B)).cuda().repeat(steps // 4 + 1, 1)[:n].contiguous()
targets = torch.empty(n, dtype=torch.float64, device="cuda")
s = torch.cuda.Stream()
teacher.eval_dev_ptr(codes.data_ptr(), n, targets.data_ptr(), s.cuda_stream)
torch.cuda.synchronize()
lam = 0.5
lam, _ = cab.train_epoch_dev(student, codes.data_ptr(), targets.data_ptr(), 8 * B, B, lam, stream=s.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
lam, st = cab.train_epoch_dev(student, codes.data_ptr(), targets.data_ptr(), n, B, lam, stream=s.cuda_stream)
e1.record(s)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(json.dumps({"slice": os.environ.get("CAB_GRAD_SLICE", "default"), "task_major": "CAB_GRAD_TASK_MAJOR" in os.environ,
                  "batch": B, "steps": steps, "ms_per_step": ms, "samples_per_s": B / ms * 1e3, "accepted": st["accepted"],
                  "E_last": st["last_new_err"] if st["last_accepted"] else st["last_err"]}))
