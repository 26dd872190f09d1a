#!/bin/bash
# ncu --set full of the walk / expand / game kernels inside one move of configs[2] (summary: profiles/r2_ncu_search.json)
out=gpurun_out
python scripts/r2_walk_target.py > $out/r2_walk_plain.log 2>&1; echo plain rc=$?
ncu --set full --clock-control none -k regex:'mcts_walk|mcts_expand|mcts_game' -s 60 -c 6 -o $out/r2_walk2 python scripts/r2_walk_target.py > $out/r2_walk2_ncu.log 2>&1; echo ncu rc=$?
ncu -i $out/r2_walk2.ncu-rep --page raw --csv > $out/r2_walk2_raw.csv 2>/dev/null
rm -f $out/r2_walk2.ncu-rep
