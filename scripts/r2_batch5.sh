#!/bin/bash
out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/r2_pytest_full4.log 2>&1; echo full rc=$?; tail -2 $out/r2_pytest_full4.log
for i in 1 2 3 4 5 6; do python -m pytest tests/test_pipeline_gpu.py -m gpu -q -k exact > $out/r2_pytest_pipe_$i.log 2>&1; echo rep $i rc=$?; done
MOVES=10 python scripts/selfplay_groups.py 1 > $out/r2_selfplay_small.jsonl 2> $out/r2_selfplay_small.err
MOVES=10 PRECISION=1 python scripts/selfplay_groups.py 1 >> $out/r2_selfplay_small.jsonl 2>> $out/r2_selfplay_small.err
cat $out/r2_selfplay_small.jsonl
