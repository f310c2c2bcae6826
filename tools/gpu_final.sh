# final check of the committed state on a B200: GPU tests, smoke(), default bench, reference arm
set -x
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log
tail -3 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_final.log; tail -2 gpurun_out/smoke_final.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'frac',d['roofline']['frac'],'slice',d['slice']['frac_of_peak'],'stem',d['stem']['value'],'clocks',d['clocks'])
r=json.loads(open('gpurun_out/bench_ref_final.json').read().strip().splitlines()[-1]); print('ref',r.get('value'),r.get('unavailable'))
PY
