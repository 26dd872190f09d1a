# forward-kernel geometry sweep: ring slot size (KB) x ring depth x host streams, through bench.py
for cfg in "48 2 4" "32 3 4" "34 3 4" "52 2 4" "40 2 4" "24 4 4" "48 2 6" "48 2 3" "48 2 8"; do
  set -- $cfg
  echo "== slot_kb=$1 stages=$2 streams=$3"
  CAB_STAGE_KB=$1 CAB_FWD_STAGES=$2 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-train --streams $3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except: print(l[:300]); continue
    r=j['roofline']; print('value %.4g e2e %.4g frac %.3f single-stream launch %.1f us parity %.2e' % (j['value'], j['e2e']['value'], r['frac'], r['launch_us_single_stream'], j['parity_max_rel_err_vs_reference_golden']))
"
done
