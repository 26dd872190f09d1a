"""One move of configs[2] (512 games x 4 roots, 800 playouts, 16 in flight) for ncu captures of the search kernels."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cab = importlib.import_module("chess-ai_b200")
w = np.fromfile(os.path.join(ROOT, "tests/golden/latest_weights.f64"), dtype="<f8")
net = cab.Network.from_weights([64, 144, 225, 100, 1], w)
sp = cab.SelfPlay(net, games=512, playouts_per_move=800, roots=4, in_flight=16, max_moves=1, network_priors=1, noise=1, save_ratio=0.1,
                  precision=int(sys.argv[1]) if len(sys.argv) > 1 else 1)
print(sp.run())
