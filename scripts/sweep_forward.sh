python -m pytest tests/test_forward_gpu.py -m gpu -x -q 2>&1 | tail -3
for cfg in "8 6 4" "8 6 8" "8 6 2" "4 3 4" "4 3 8" "4 8 4" "8 4 4"; do
  set -- $cfg
  echo "== warps=$1 stages=$2 streams=$3"
  CAB_FWD_WARPS=$1 CAB_FWD_STAGES=$2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --streams $3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except: print(l[:300]); continue
    r=j['roofline']; print('value %.4g e2e %.4g frac %.3f single-stream launch %.1f us parity %.2e' % (j['value'], j['e2e']['value'], r['frac'], r['launch_us_single_stream'], j['parity_max_rel_err_vs_reference_golden']))
"
done
