#!/bin/bash
# GPU trip: bench (ours + reference), launch list and one full ncu capture of the sweeps
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "ours rc=$?"; tail -c 4000 gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 1500 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_multiply_rows|k_propagate_cols|k_potential_cols|k_transmit_rows|k_bandlimit_cols|k_density_rows' -s 60 -c 8 -o gpurun_out/prof_sweeps $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
