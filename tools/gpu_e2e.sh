set -x
FDES_B200_TIMING=1 python bench.py --steps 20 --warmup 3 --no-cpu --no-stem > gpurun_out/bench_pinned.json 2> gpurun_out/bench_pinned.err; echo "rc=$?"
grep "fdes_b200 timing" gpurun_out/bench_pinned.err | tail -4
grep "\[engine\]" gpurun_out/bench_pinned.err | tail -12
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_pinned.json').read().strip().splitlines()[-1]); print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'])
PY
