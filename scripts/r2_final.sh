#!/bin/bash
# final evidence of the round at N = 1: GPU tests, the default bench line, the reference arm, the launch list of the bench
out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/r2_pytest_final.log 2>&1; echo pytest rc=$?; tail -1 $out/r2_pytest_final.log
python bench.py > $out/r2_bench_final_n1.json 2> $out/r2_bench_final_n1.err; echo bench rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > $out/r2_bench_final_ref.json 2> $out/r2_bench_final_ref.err; echo ref rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/r2_bench_launches_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-wide --selfplay-moves 1 > $out/r2_bl_final.log 2>&1; echo ncu rc=$?
