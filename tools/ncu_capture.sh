#!/bin/bash
# One gpurun call that refreshes the profiling evidence for the current kernels (run from the repo root ON THE GPU BOX):
#
#   gpurun --timeout 1200 -- 'bash tools/ncu_capture.sh r2a'
#
# Writes into gpurun_out/ (copy what you want judged into profiles/; gpurun brings back at most 64 MiB, so the
# .ncu-rep files are reduced to CSV on the box and only the leaf-sponge and contiguous-NTT reports are kept):
#   launches_<tag>.csv              every launch of a short run with its device time (serialised, cold cache: compare shares)
#   prof_<tag>_<kernel>_raw.csv     ncu --set full, raw page, one launch of each dominant kernel after the warm-up proofs:
#                                   dram bytes, pipe utilisation (alu / fma / fmaheavy), issue slots, stall reasons, registers, occupancy
#   prof_<tag>_<kernel>_src.csv     source page (per-line instruction counts and stall samples; needs -lineinfo), top lines
#   bench_<tag>.json                the plain bench line of the same build (never taken under the profiler)
# tools/profile_prove.py keeps the traces device-resident and proves --proves times (2 warm-up + 1 profiled).
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
PY="python tools/profile_prove.py --log-height 20 --proves 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${TAG}.csv $PY > gpurun_out/launches_${TAG}.log 2>&1
# kernel:launches to skip (the first two proofs' launches of that kernel)
for K in k_leaf_hash:6 k_compress:348 k_fwd_contig:226 k_fwd_strided:226 k_intt_strided:16 k_intt_contig:16 k_deep:2 k_fri_leaf:14 k_constraints:6 k_ood_dot:20; do
    NAME=${K%%:*}; SKIP=${K##*:}
    REP=gpurun_out/prof_${TAG}_${NAME}
    ncu --set full --clock-control none --import-source on -k regex:${NAME} -s ${SKIP} -c 1 -f -o ${REP} $PY > ${REP}.log 2>&1
    ncu -i ${REP}.ncu-rep --page raw --csv > ${REP}_raw.csv 2>/dev/null
    ncu -i ${REP}.ncu-rep --page source --csv 2>/dev/null | head -4000 > ${REP}_src.csv
    case ${NAME} in k_leaf_hash|k_fwd_contig) ;; *) rm -f ${REP}.ncu-rep ;; esac
done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.log
tail -c 600 gpurun_out/bench_${TAG}.json
du -sh gpurun_out
