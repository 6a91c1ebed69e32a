#!/bin/bash
# One profiling round on the GPU box (run under gpurun):  tools/profile_round.sh <tag> [workload]
#   1. the plain bench must exit 0 first (numbers are never taken under ncu)
#   2. ncu launch list of the same command (durations per launch: cold-cache, serialised -> compare SHARES)
#   3. ncu --set full of one step's kernels from the middle of the same command, raw page exported for
#      tools/extract_ncu_summary.py (the report itself is kept only if it is small enough to travel back)
#   4. ncu --set full with sources of ONE pair-kernel launch (hot-instruction table)
set -u
TAG=${1:-r02_v1}; WL=${2:-c3}
CMD="python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-extra"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain bench failed"; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_l.log 2>&1
python tools/summarise_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launch_summary.csv
head -45 gpurun_out/${TAG}_launch_summary.csv
NK=$(grep -c -E '"(void )?k_' gpurun_out/${TAG}_launches.csv)
SKIP=$((NK / 2))
echo "kernel launches in the command: $NK, full capture skips $SKIP"
ncu --set full --clock-control none -k regex:"^k_|k_fft16|k_pair" --launch-skip $SKIP -c ${COUNT:-45} -f -o gpurun_out/prof_${TAG}_full $CMD > gpurun_out/${TAG}_ncu_f.log 2>&1
ncu -i gpurun_out/prof_${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ls -la gpurun_out/prof_${TAG}_full.ncu-rep gpurun_out/${TAG}_full_raw.csv
[ $(stat -c %s gpurun_out/prof_${TAG}_full.ncu-rep) -gt 30000000 ] && rm -f gpurun_out/prof_${TAG}_full.ncu-rep
ncu --set full --import-source on --clock-control none -k regex:k_pair_tiles --launch-skip 8 -c 1 -f -o gpurun_out/prof_${TAG}_pair $CMD > gpurun_out/${TAG}_ncu_p.log 2>&1
ncu -i gpurun_out/prof_${TAG}_pair.ncu-rep --page source --csv > gpurun_out/${TAG}_pair_src.csv 2>/dev/null
ls -la gpurun_out/prof_${TAG}_pair.ncu-rep
du -sh gpurun_out
