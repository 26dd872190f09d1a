"""configs[2] self-play as G independent search groups of 512 / G games, each driven by its own host thread on its own
streams (cab_mcts_run releases the GIL): one group's tree walk fills the SMs the other group's evaluator leaves idle.
usage: python scripts/selfplay_groups.py [groups ...]   (default: 1 2 4)"""
import importlib, json, os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cab = importlib.import_module("chess-ai_b200")
weights = np.fromfile(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "latest_weights.f64"), dtype="<f8")
net = cab.Network.from_weights([64, 144, 225, 100, 1], weights)
moves = int(os.environ.get("MOVES", "10"))
prec = int(os.environ.get("PRECISION", "0"))
roots = int(os.environ.get("ROOTS", "8"))
infl = int(os.environ.get("IN_FLIGHT", "16"))
for G in [int(a) for a in sys.argv[1:]] or [1, 2, 4]:
    sps = [cab.SelfPlay(net, games=512 // G, playouts_per_move=800, roots=roots, in_flight=infl, max_moves=moves, network_priors=1, noise=1,
                        save_ratio=0.1, seed=1234 + 1000 * g, precision=prec) for g in range(G)]
    stats = [None] * G
    def work(g):
        stats[g] = sps[g].run()
    ths = [threading.Thread(target=work, args=(g,)) for g in range(G)]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    secs = time.perf_counter() - t0
    tot = lambda k: sum(s[k] for s in stats)
    print(json.dumps({"groups": G, "games": 512, "moves": moves, "precision": prec, "roots": roots, "in_flight": infl, "seconds": secs, "playouts_per_s": tot("playouts") / secs,
                      "positions_per_s": tot("boards_evaluated") / secs, "moves_played": tot("moves_played"),
                      "evaluator_seconds_sum": tot("evaluator_seconds"), "rounds": [s["rounds"] for s in stats],
                      "eval_full": tot("eval_full"), "arena_full": tot("arena_full")}), flush=True)
    for sp in sps: sp.close()
