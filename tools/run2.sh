P=$PWD/summationbypartsoperatorsextra.jl_b200
for rep in 1 2; do
for v in old noping new; do
  lib=$P/libsbp_b200_$v.so; [ $v = new ] && lib=$P/libsbp_b200.so
  for wl in "banded --n 104" "banded --n 101" "config1b --n 104" config1b config3; do SBP_B200_LIB=$lib python tools/time_config.py --workload $wl --steps 40 --tag $v 2>/dev/null; done
done; done | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(f\"{d['tag']:8s} {d['workload']:9s} N={d['N']:5d} {d['ms_median']:.4f} ms {d['gdof_s']:.1f} GDOF/s\")" | sort
