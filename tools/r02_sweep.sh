#!/bin/bash
# GPU session helper: L2-policy sweep of the tcgen05 contraction at config 4 (one GPU).  Usage: tools/r02_sweep.sh TAG
TAG=${1:-r02}
OUT=gpurun_out
run() {  # name, extra args
  name=$1; shift
  python bench.py --steps 5 --warmup 2 --no-sub-records --no-cpu-baseline --no-parity "$@" > $OUT/${TAG}_sweep_${name}.json 2> $OUT/${TAG}_sweep_${name}.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_sweep_${name}.json").read().strip().splitlines()[-1])
    s = d["roofline"]["stage_ms"]
    print("${name}: step %.2f ms, contract %.2f, softmax %.2f, lp %.2f, dedup %.2f, other %.2f, e2e %.2f ms, clocks %s" % (
        d["ms_per_step"], s["contract"], s["softmax"], s["lp_table"], s["dedup"], s["other"], d["e2e"]["ms_per_step"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("${name}: FAILED", e)
PY
}
run base
run pm --opt item_order=1
run pm_h8 --opt item_order=1 --opt hot_rank=8
run pm_h16 --opt item_order=1 --opt hot_rank=16
run pm_h32 --opt item_order=1 --opt hot_rank=32
run pm_h64 --opt item_order=1 --opt hot_rank=64
run pm_h16n --opt item_order=1 --opt hot_rank=16 --opt cold_evict_first=0
run gm_h4 --opt hot_rank=4
run gm_h8 --opt hot_rank=8
