set -x
mkdir -p gpurun_out/l2
for i in 1 2; do
python bench.py --steps 30 --warmup 3 --no-cpu --no-stem > gpurun_out/l2/off_$i.json 2> gpurun_out/l2/off_$i.err
FDES_B200_L2_PERSIST=1 FDES_B200_TIMING=1 python bench.py --steps 30 --warmup 3 --no-cpu --no-stem > gpurun_out/l2/on_$i.json 2> gpurun_out/l2/on_$i.err
done
grep -h "L2 persistence" gpurun_out/l2/on_1.err | head -2
FDES_B200_L2_PERSIST=1 python bench.py --steps 30 --warmup 3 --no-cpu --no-stem --batch 4 > gpurun_out/l2/on_b4.json 2> gpurun_out/l2/on_b4.err
python bench.py --steps 30 --warmup 3 --no-cpu --no-stem --batch 4 > gpurun_out/l2/off_b4.json 2> gpurun_out/l2/off_b4.err
python -m pytest tests/test_qsc.py -q -m gpu -k "cli_stem or stem_scan" > gpurun_out/pytest_stem.log 2>&1; tail -3 gpurun_out/pytest_stem.log
for f in gpurun_out/l2/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['sweeps'].items()})
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
