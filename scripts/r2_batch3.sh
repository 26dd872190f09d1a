#!/bin/bash
out=gpurun_out
python -m pytest tests/test_mcts_gpu.py tests/test_chess_rules.py tests/test_leafq_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q > $out/r2_pytest_mcts2.log 2>&1; echo pytest rc=$?; tail -3 $out/r2_pytest_mcts2.log
MOVES=10 python scripts/selfplay_groups.py 1 > $out/r2_selfplay_noinline.jsonl 2> $out/r2_selfplay_noinline.err
MOVES=10 PRECISION=1 python scripts/selfplay_groups.py 1 >> $out/r2_selfplay_noinline.jsonl 2>> $out/r2_selfplay_noinline.err
cat $out/r2_selfplay_noinline.jsonl; tail -2 $out/r2_selfplay_noinline.err
