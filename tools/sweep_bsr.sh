#!/bin/bash
# bsr kernel variant sweep on the GPU box: libs built by tools/mkvar.sh; results -> gpurun_out/sweep_bsr.jsonl
OUT=gpurun_out/sweep_bsr.jsonl
: > $OUT
P=summationbypartsoperatorsextra.jl_b200
run() { # tag lib env...
  local tag=$1 lib=$2; shift 2
  for wl in config3 config1b "banded --n 104"; do
    env "$@" SBP_B200_LIB=$PWD/$P/$lib python tools/time_config.py --workload $wl --steps 30 --tag "$tag" >> $OUT 2>> gpurun_out/sweep_bsr.err
  done
}
run ping libsbp_b200.so X=1
run noping libsbp_b200_noping.so X=1
run ch12 libsbp_b200_ch12.so X=1
for mt in 5 6; do run ch12_mt$mt libsbp_b200_ch12.so SBP_BSR_MT=$mt; run ping_mt$mt libsbp_b200.so SBP_BSR_MT=$mt; done
cat $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(f\"{d['tag']:14s} {d['workload']:9s} N={d['N']} {d['ms_median']:.4f} ms {d['gdof_s']:.1f} GDOF/s\")"
