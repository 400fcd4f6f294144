for n in 32 40 64 65 101 104 128; do for d in "" 1; do
  SBP_BENCH_DENSE=$d python tools/time_config.py --workload banded --n $n --steps 30 --tag "n${n}_dense${d}" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:160]); continue
    print(d['tag'], d['kernel'], round(d['ms_median'],4), 'ms', round(d['gdof_s'],1), 'GDOF/s')"
done; done
