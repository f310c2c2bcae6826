#!/bin/bash
# GPU trip: bench (ours + reference), launch list and one full ncu capture of the top kernel
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "ours rc=$?"; tail -c 3000 gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
python bench.py --steps 5 --warmup 3 --batch 4 --no-cpu > gpurun_out/bench_ours_b4.json 2> gpurun_out/bench_ours_b4.err; echo "ours b4 rc=$?"; tail -c 3000 gpurun_out/bench_ours_b4.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 2000 gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --configs-per-step 16"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_multiply_rows|k_propagate_cols|k_potential_cols|k_transmit_rows|k_bandlimit_cols|k_density_rows' -s 600 -c 6 -o gpurun_out/prof_r1 $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
