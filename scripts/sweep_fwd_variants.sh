#!/bin/bash
# forward-kernel variant sweep through the native harness (one process per variant: the library reads CAB_FWD_* once)
#   scripts/sweep_fwd_variants.sh > gpurun_out/fwd_sweep.jsonl
B=chess-ai_b200/lib/fwd_bench
W=tests/golden/latest_weights.f64
run() { env "$@" $B $W 1048576 4096 ${STREAMS:-4} 6; }
run CAB_FWD_FUSE_OUT=1
run CAB_FWD_FUSE_OUT=0
