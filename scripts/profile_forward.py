"""Debug helper: per-warp phase cycle counters of single forward launches (CAB_FWD_PROFILE=1 must be set)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cab = importlib.import_module("chess-ai_b200")
w = np.fromfile("tests/golden/latest_weights.f64", dtype="<f8")
net = cab.Network.from_weights([64, 144, 225, 100, 1], w)
codes = cab.synthetic_boards(65536)
for n in (4096, 4096, 9472, 65536):
    net.predict_batch(codes[:n])
