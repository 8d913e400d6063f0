#!/bin/bash
# Profiling pass of one round on the GPU box (run through gpurun, one GPU):
#   bash tools/profile_round.sh r01final          (KERNELS="name:skip ..." and SKIP_LIST=1 narrow the pass)
# 1. plain bench (must exit 0 before anything runs under ncu), 2. launch list of the same command,
# 3. one `ncu --set full` capture per hot kernel.  Everything lands in gpurun_out/; tools/ncu_excerpt.py and
# tools/launch_shares.py turn the captures into the summaries kept under profiles/.
tag=${1:-r01final}
out=gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity"
$B > $out/${tag}_plain.log 2>&1 || { echo "plain bench failed"; tail -5 $out/${tag}_plain.log; exit 1; }
[ -n "$SKIP_LIST" ] || ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file $out/${tag}_launches.csv $B > $out/${tag}_launches.log 2>&1
# kernel:launches to skip = the warm-up step's launches of the resident-counts job (per-joint kernels launch twice a
# step, the front kernels once; the chunked one-shot call that follows launches the front kernels per chunk)
for ks in ${KERNELS:-contract_i8_kernel:2 lp_rows_q_kernel:1 softmax_i8_warp_kernel:2 dedup_bitmap_emit_kernel:1 \
          ratio_summary_kernel:1 sentinel_range_kernel:2}; do
    k=${ks%%:*}
    ncu --set full --clock-control none --import-source on -k regex:"^${k}" -s ${ks##*:} -c 1 -f -o $out/${tag}_${k} $B \
        > $out/${tag}_${k}.log 2>&1
    ls -la $out/${tag}_${k}.ncu-rep
done
