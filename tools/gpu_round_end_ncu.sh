#!/bin/bash
# Round-end GPU trip, part 2: ncu launch list and full captures of the six sweeps at 1024^2 and 2048^2
# (each only after the same command has exited 0 without ncu).  Summaries: tools/summarize_profiles.py.
tag=${1:-rX}; out=gpurun_out/$tag; mkdir -p $out
K='k_multiply_rows|k_propagate_cols|k_potential_cols|k_transmit_rows|k_bandlimit_cols|k_density_rows'
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-stem --no-job --configs-per-step 20"
$CMD > $out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 500 --csv --log-file $out/launches.csv $CMD > $out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$CMD > $out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$K" -s 60 -c 8 -o $out/prof_sweeps $CMD > $out/ncu_full.log 2>&1
echo "ncu full 1024 rc=$?"; tail -1 $out/ncu_full.log
CMD2="$CMD --workload au_2048"
$CMD2 > $out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$K" -s 60 -c 8 -o $out/prof_sweeps_2048 $CMD2 > $out/ncu_full_2048.log 2>&1
echo "ncu full 2048 rc=$?"; tail -1 $out/ncu_full_2048.log
ls -la $out
