python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in -2 0; do SBP_GEMM_VARIANT=$v python tools/time_config.py --workload config5 --steps 10 --tag gemm$v 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['tag'], d['ms_median'], 'ms', 2*d['N']*d['N']*d['B']/d['ms_median']/1e9, 'TFLOP/s')"; done
