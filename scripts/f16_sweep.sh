#!/bin/bash
# A/B of the fused fp16 kernel's chunk sizes (CAB_F16_CH) and the round-1 serial kernel (CAB_F16_SERIAL=1)
out=gpurun_out/${1:-r2_f16_sweep}.jsonl
rm -f $out
for ch in ${2:-16 32 48}; do CAB_F16_CH=$ch timeout 100 python scripts/f16_ab.py >> $out 2>>${out%.jsonl}.err; echo "ch=$ch rc=$?"; done
python - <<PY
import json
for l in open("$out"):
    d=json.loads(l); print({k:(round(v,6) if "diff" in k else "%.3e"%v) for k,v in d.items() if k.startswith("evals") or k.startswith("batches") or k in ("latest_maxdiff_5000","small_maxdiff_5000","small_maxdiff_129","small_maxdiff_1")})
PY
tail -3 ${out%.jsonl}.err
