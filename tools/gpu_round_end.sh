#!/bin/bash
# Round-end GPU trip, part 1 (gpurun --timeout 1500 -- 'bash tools/gpu_round_end.sh <tag>'): the whole GPU
# test suite, smoke(), the bench lines of both arms and of the large grids, the column microbenchmark.
tag=${1:-rX}; out=gpurun_out/$tag; mkdir -p $out
python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > $out/bench_ours.json 2> $out/bench_ours.err; echo "ours rc=$?"
for wl in au_2048 slab_4096 slab_4096_full; do
  python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu --no-stem --no-job > $out/bench_$wl.json 2> $out/bench_$wl.err; echo "$wl rc=$?"
done
tools/microbench/col_bench > $out/col_bench.txt 2>&1
python - $out <<'PY'
import json, sys, pathlib
for f in sorted(pathlib.Path(sys.argv[1]).glob("bench_*.json")):
    try:
        d = json.loads(f.read_text().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(f.name, d.get("impl", "ours"), round(d["value"]), "e2e", round(d["e2e"]["value"]), "slice", r.get("frac_of_8TBps"),
              "step", r.get("whole_step", {}).get("frac_of_8TBps"), {k[:2]: round(v["ms"] * 1e3, 1) for k, v in d.get("sweeps", {}).items()},
              "stem", (d.get("stem") or {}).get("value"))
    except Exception as e:
        print(f.name, "unreadable", e)
PY
