#!/bin/bash
# FRAP kernel evidence (one gpurun call): agent parity tests, the c4 bench line, one full ncu capture of k_policy_frap
O=gpurun_out; T=${TAG:-r02x}
python -m pytest tests/test_agents.py -m gpu -q 2>&1 | tail -2
python bench.py --config c4 --steps 40 --warmup 5 > $O/${T}_bench_c4.json 2> $O/${T}_c4.err; cut -c1-200 $O/${T}_bench_c4.json
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_policy_frap -s 2 -c 1 -f -o $O/${T}_c4_frap \
  python bench.py --config c4 --steps 8 --warmup 3 --no-cpu --profile > $O/${T}_c4_frap_ncu.log 2>&1
ncu -i $O/${T}_c4_frap.ncu-rep --page raw --csv > $O/${T}_c4_frap_raw.csv 2>/dev/null
rm -f $O/${T}_c4_frap.ncu-rep
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/${T}_c4_launches.csv \
  python bench.py --config c4 --steps 8 --warmup 3 --no-cpu --profile > $O/${T}_c4_launch_bench.log 2>&1
python tools/launch_shares.py $O/${T}_c4_launches.csv
python tools/ncu_summary.py $O/${T}_c4_frap_raw.csv | grep -E "time_duration|issue_active|bank_conflicts|short_scoreboard|registers"
