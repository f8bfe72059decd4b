#!/bin/bash
# One gpurun call that refreshes the profiling evidence for the current kernels (run from the repo root ON THE GPU BOX):
#
#   gpurun --timeout 900 -- 'bash tools/ncu_capture.sh r2a'
#
# Writes into gpurun_out/ (copy what you want judged into profiles/):
#   launches_<tag>.csv         every launch of a short bench run with its device time (serialised, cold cache: compare shares)
#   prof_<tag>_<kernel>.ncu-rep + *_raw.csv   ncu --set full of the dominant kernels (3 launches each, after the warm-up proofs):
#                              dram bytes, pipe utilisation (alu / fma / fmaheavy), issue-slot utilisation, stall reasons, registers
#   bench_<tag>.json           the plain bench line of the same build (never taken under the profiler)
# tools/profile_prove.py keeps the traces device-resident and proves --proves times (2 warm-up + 1 profiled).
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
PY="python tools/profile_prove.py --log-height 20 --proves 3"
# kernels per proof (profiles/launches_r1i_summary.md): skip the first two proofs' launches of each kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${TAG}.csv $PY > gpurun_out/launches_${TAG}.log 2>&1
for K in k_leaf_hash:6 k_compress:348 k_fwd_contig:226 k_fwd_strided:226 k_intt_strided:16 k_intt_contig:16 k_deep:2 k_fri_leaf:14 k_constraints:6 k_ood_dot:20; do
    NAME=${K%%:*}; SKIP=${K##*:}
    ncu --set full --clock-control none --import-source on -k regex:${NAME} -s ${SKIP} -c 3 -f -o gpurun_out/prof_${TAG}_${NAME} $PY > gpurun_out/prof_${TAG}_${NAME}.log 2>&1
    ncu -i gpurun_out/prof_${TAG}_${NAME}.ncu-rep --page raw --csv 2>/dev/null | grep -E 'Metric Name|gpu__time_duration.sum|dram__bytes_(read|write)\.sum|gpu__dram_throughput|sm__inst_executed\.sum|sm__inst_executed_pipe_(alu|fma|fmaheavy|fmalite)\.sum|sm__pipe_(alu|fma|fmaheavy)_cycles_active|sm__issue_active|smsp__issue_active|smsp__warp_issue_stalled|smsp__average_warp|sm__warps_active|launch__registers_per_thread|launch__occupancy' > gpurun_out/prof_${TAG}_${NAME}_raw.csv
done
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.log
tail -c 600 gpurun_out/bench_${TAG}.json
