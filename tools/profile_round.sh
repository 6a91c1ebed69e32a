#!/bin/bash
# One profiling round on the GPU box (run under gpurun):  tools/profile_round.sh <tag> [workload]
#   1. the plain bench must exit 0 first (numbers are never taken under ncu)
#   2. ncu launch list of the same command (durations per launch: cold-cache, serialised -> compare SHARES)
#   3. ncu --set full of the step's kernels (a few launches each), raw page exported for tools/extract_ncu_summary.py
set -u
TAG=${1:-r01_v3}; WL=${2:-c3}
CMD="python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain bench failed"; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
if [ -z "${SKIP_LIST:-}" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_l.log 2>&1
python tools/summarise_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launch_summary.csv
head -40 gpurun_out/${TAG}_launch_summary.csv
fi
# full set: skip the set-up launches, then 2 steps' worth of kernels
# full set: one step's worth of kernels after the set-up launches (a report of 90 launches with sources is > 64 MiB and
# does not travel back: keep the count low; the raw page is exported on the box, the report kept only if small)
ncu --set full --clock-control none -k regex:"^k_|k_fft16|k_pair" --launch-skip ${SKIP:-430} -c ${COUNT:-36} -f -o gpurun_out/prof_${TAG}_full $CMD > gpurun_out/${TAG}_ncu_f.log 2>&1
ncu -i gpurun_out/prof_${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ls -la gpurun_out/prof_${TAG}_full.ncu-rep gpurun_out/${TAG}_full_raw.csv
[ $(stat -c %s gpurun_out/prof_${TAG}_full.ncu-rep) -gt 40000000 ] && rm -f gpurun_out/prof_${TAG}_full.ncu-rep
