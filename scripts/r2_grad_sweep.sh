#!/bin/bash
# A/B of the gradient kernel's item order and slice size + DRAM traffic of grad_kernel / encode_kernel (ncu).
out=gpurun_out
python -m pytest tests/test_forward_gpu.py tests/test_train_gpu.py -m gpu -x -q > $out/r2_pytest_enc_grad.log 2>&1; echo pytest rc=$?
: > $out/r2_grad_ab.jsonl
for s in 128 256 512 1024; do CAB_GRAD_SLICE=$s python scripts/grad_ab.py 32 >> $out/r2_grad_ab.jsonl 2>> $out/r2_grad_ab.err; done
CAB_GRAD_SLICE=640 CAB_GRAD_TASK_MAJOR=1 python scripts/grad_ab.py 32 >> $out/r2_grad_ab.jsonl 2>> $out/r2_grad_ab.err
python scripts/grad_ab.py 32 4096 >> $out/r2_grad_ab.jsonl 2>> $out/r2_grad_ab.err
CAB_GRAD_SLICE=640 CAB_GRAD_TASK_MAJOR=1 python scripts/grad_ab.py 32 4096 >> $out/r2_grad_ab.jsonl 2>> $out/r2_grad_ab.err
cat $out/r2_grad_ab.jsonl
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
    -k regex:grad_kernel -s 8 -c 3 --csv --log-file $out/r2_grad_traffic.csv python scripts/grad_ab.py 4 > $out/r2_grad_traffic.log 2>&1; echo ncu rc=$?
CAB_GRAD_SLICE=640 CAB_GRAD_TASK_MAJOR=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
    -k regex:grad_kernel -s 8 -c 3 --csv --log-file $out/r2_grad_traffic_taskmajor.csv python scripts/grad_ab.py 4 > $out/r2_grad_traffic_tm.log 2>&1; echo ncu rc=$?
cat > /tmp/enc.py <<'PY'
import importlib, torch
cab = importlib.import_module("chess-ai_b200")
n = 1 << 21
r = torch.randint(0, 7, (n, 64, 3), dtype=torch.uint8, device="cuda")
o = torch.empty(n, 64, dtype=torch.int8, device="cuda")
for _ in range(4):
    cab._ck(cab.lib().cab_encode_dev(r.data_ptr(), n, o.data_ptr(), None))
torch.cuda.synchronize()
PY
PYTHONPATH=. ncu --set full --clock-control none -k regex:encode_kernel -s 2 -c 1 -o $out/r2_encode_full python /tmp/enc.py > $out/r2_encode_ncu.log 2>&1; echo ncu rc=$?
ncu -i $out/r2_encode_full.ncu-rep --page raw --csv > $out/r2_encode_full_raw.csv 2>/dev/null
python -m pytest tests/test_dropin_gpu.py -m gpu -x -q > $out/r2_pytest_dropin.log 2>&1; echo pytest dropin rc=$?
MOVES=10 python scripts/selfplay_groups.py 1 2 4 > $out/r2_selfplay_groups.jsonl 2> $out/r2_selfplay_groups.err; cat $out/r2_selfplay_groups.jsonl
