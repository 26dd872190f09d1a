#!/bin/bash
# passes per CTA of overlapping 4 096-board forward launches x streams: bench.py's eval legs only. RECORD of an experiment:
# the CAB_FWD_PASSES knob it drove was removed again (launch_forward, cab_api.cu: one pass per CTA measured best,
# profiles/r2_fwd_passes.jsonl); with the shipped library every line of this sweep is the 1-pass geometry.
out=gpurun_out
: > $out/r2_fwd_passes.jsonl
run() { # passes streams
  CAB_FWD_PASSES=$1 python bench.py --streams $2 --no-train --no-selfplay --no-wide --no-cpu-baseline 2>> $out/r2_fwd_passes.err | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'passes': $1, 'streams': $2, 'value': d['value'], 'frac': d['roofline']['frac'], 'peak': d['roofline']['peak'], 'e2e': d['e2e']['value'], 'sm_mhz': d['clocks']['sm_mhz']}))" >> $out/r2_fwd_passes.jsonl
}
run 1 4; run 1 8; run 2 6; run 2 8; run 4 12; run 4 16; run 2 12
cat $out/r2_fwd_passes.jsonl
