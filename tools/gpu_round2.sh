set -x
mkdir -p gpurun_out/golden gpurun_out/r1v4
python tools/make_golden.py --out gpurun_out/golden qsc64.qsc qsctilt64.qsc > gpurun_out/golden_qsc.log 2>&1
cp gpurun_out/golden/qsc*.npz gpurun_out/golden/qsc*.txt tests/golden/
python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "srtio3_800|ours-vs|passed|failed|rc=" gpurun_out/pytest_gpu.log | tail -12
python -m pytest tests -m "not gpu" -q > gpurun_out/pytest_cpu.log 2>&1; tail -2 gpurun_out/pytest_cpu.log
python bench.py --steps 20 --warmup 3 --no-cpu --no-stem > gpurun_out/r1v4/bench_b8.json 2> gpurun_out/r1v4/bench_b8.err
python bench.py --steps 20 --warmup 3 --no-cpu --no-stem --batch 16 --configs-per-step 16 > gpurun_out/r1v4/bench_b16.json 2> gpurun_out/r1v4/bench_b16.err
python bench.py --steps 20 --warmup 3 --no-cpu --no-stem --batch 16 --configs-per-step 32 > gpurun_out/r1v4/bench_b16c32.json 2> gpurun_out/r1v4/bench_b16c32.err
python bench.py --steps 20 --warmup 3 --no-cpu --no-stem --batch 32 --configs-per-step 32 > gpurun_out/r1v4/bench_b32.json 2> gpurun_out/r1v4/bench_b32.err
for f in gpurun_out/r1v4/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['slice'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
