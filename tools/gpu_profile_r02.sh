#!/bin/bash
# round-2 profiling evidence (one gpurun call): for every bench config the ncu launch list of the bench command
# (per-launch durations: the kernel's SHARE of the step) and one full capture of the fused env-step kernel.
#   usage: tools/gpu_profile_r02.sh [configs...]      TAG=r02x
cd "$(dirname "$0")/.."
T=${TAG:-r02}
O=gpurun_out
for c in ${@:-c2 c3 c5}; do
  # (--profile: cudaProfilerStart at the first TIMED step, after the config's own pre-roll: the loaded network)
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/${T}_${c}_launches.csv \
      python bench.py --config $c --steps 8 --warmup 3 --no-cpu --profile > $O/${T}_${c}_launch_bench.log 2>&1
  # the fused env step of the second timed step (c3 / c4 launch k_run twice per step: fast pass + overflow pass)
  SKIP=1; if [ $c = c3 ] || [ $c = c4 ]; then SKIP=2; fi
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_run -s $SKIP -c 1 -f -o $O/${T}_${c}_krun \
      python bench.py --config $c --steps 8 --warmup 3 --no-cpu --profile > $O/${T}_${c}_ncu.log 2>&1
  # gpurun merges at most 64 MiB back: keep the exports (raw page, per-instruction source page), not the 23 MB reports
  ncu -i $O/${T}_${c}_krun.ncu-rep --page raw --csv > $O/${T}_${c}_raw.csv 2>/dev/null
  ncu -i $O/${T}_${c}_krun.ncu-rep --page source --csv > $O/${T}_${c}_src.csv 2>/dev/null
  rm -f $O/${T}_${c}_krun.ncu-rep
  ls -la $O/${T}_${c}_raw.csv $O/${T}_${c}_src.csv
done
