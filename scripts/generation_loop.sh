#!/bin/bash
# Self-play -> packed cases -> trainer -> checkpoint rotation -> self-play with the new network, G generations.
# The whole N1-N4 pipeline on one GPU (chess-ai_b200/lib/selfplay + lib/train). Prints the tools' JSON lines.
#   scripts/generation_loop.sh [generations] [games] [playouts] [moves] [train_steps] [batch] [work_dir]
set -e
G=${1:-3}; GAMES=${2:-512}; PLAYOUTS=${3:-200}; MOVES=${4:-30}; STEPS=${5:-300}; BATCH=${6:-4096}; WORK=${7:-/tmp/cab_generations}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
L=$ROOT/chess-ai_b200/lib
rm -rf "$WORK"; mkdir -p "$WORK/network_dump"
python - "$ROOT" "$WORK" <<'PY'
import importlib, sys, numpy as np
sys.path.insert(0, sys.argv[1])
cab = importlib.import_module("chess-ai_b200")
w = np.fromfile(sys.argv[1] + "/tests/golden/latest_weights.f64", dtype="<f8")
cab.Network.from_weights([64, 144, 225, 100, 1], w).dump(sys.argv[2] + "/start.txt")
PY
NET=$WORK/start.txt
for g in $(seq 1 $G); do
  SHARD=$WORK/gen$g.shard
  $L/selfplay $NET $GAMES $PLAYOUTS $MOVES - 8 1 0 $((1000 + g)) $SHARD 1.0 fp64 4
  $L/train $NET $SHARD $WORK/network_dump batch $STEPS $BATCH 0 $g 100
  NET=$WORK/network_dump/latest.txt
done
ls $WORK/network_dump | tr '\n' ' '; echo
