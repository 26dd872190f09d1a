#!/bin/bash
out=gpurun_out
python -m pytest tests/test_forward_gpu.py -m gpu -x -q > $out/r2_pytest_enc2.log 2>&1; echo pytest rc=$?
MOVES=10 python scripts/selfplay_groups.py 1 2 4 > $out/r2_selfplay_groups.jsonl 2> $out/r2_selfplay_groups.err; cat $out/r2_selfplay_groups.jsonl; tail -2 $out/r2_selfplay_groups.err
bash scripts/r2_fwd_passes.sh
python bench.py --no-train --no-selfplay --no-cpu-baseline 2> $out/r2_bench_enc.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps(d['encoder'])); print(json.dumps(d['wide_bf16']))" | tee $out/r2_bench_enc.json
