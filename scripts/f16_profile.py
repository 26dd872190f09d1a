"""CAB_F16_PROFILE=1: clock stamps of the fused fp16 kernel's second tile per CTA (one launch of 2^16 boards)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cab = importlib.import_module("chess-ai_b200")
w = np.fromfile(os.path.join(ROOT, "tests/golden/latest_weights.f64"), dtype="<f8")
net = cab.Network.from_weights([64, 144, 225, 100, 1], w)
codes = cab.synthetic_boards(1 << 16)
for _ in range(2):
    net.predict_batch_f16(codes)
