#!/bin/bash
# search-kernel check after a change: the MCTS / rules / leaf-queue / pipeline GPU tests, then configs[2] over 10 moves with
# the fp64 and the fp16 evaluator (scripts/selfplay_groups.py 1)
out=gpurun_out
python -m pytest tests/test_mcts_gpu.py tests/test_chess_rules.py tests/test_leafq_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q > $out/r2_pytest_mcts2.log 2>&1; echo pytest rc=$?; tail -3 $out/r2_pytest_mcts2.log
MOVES=10 python scripts/selfplay_groups.py 1 > $out/r2_selfplay_check.jsonl 2> $out/r2_selfplay_check.err
MOVES=10 PRECISION=1 python scripts/selfplay_groups.py 1 >> $out/r2_selfplay_check.jsonl 2>> $out/r2_selfplay_check.err
cat $out/r2_selfplay_check.jsonl; tail -2 $out/r2_selfplay_check.err
