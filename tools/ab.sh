#!/bin/bash
# A/B of library variants built with `make OBJ=../build_<v> OUT=../lib_<v> BIN=../bin_<v> EXTRA=-D...`:
#   tools/ab.sh <outdir> <workload> <variant> [<variant> ...]     ("base" = fdes_b200/lib)
out=$1; wl=$2; shift 2
mkdir -p $out
for v in "$@"; do
  lib=fdes_b200/lib_$v/libfdes_b200.so; [ "$v" = base ] && lib=fdes_b200/lib/libfdes_b200.so
  FDES_B200_LIB=$PWD/$lib python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu --no-stem --no-job > $out/${wl}_$v.json 2> $out/${wl}_$v.err
  python - "$out/${wl}_$v.json" "$v" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], d["config"]["workload"], round(d["value"]), d["roofline"]["frac_of_8TBps"], d["roofline"]["whole_step"]["frac_of_8TBps"],
      {k[:2]: round(v["ms"] * 1e3, 1) for k, v in d["sweeps"].items()})
PY
done
