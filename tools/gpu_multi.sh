#!/bin/bash
# Multi-GPU trip (gpurun --gpus N -- 'bash tools/gpu_multi.sh N'): multi-GPU tests, the torchrun bench line
# (weak scaling + strong-scaling jobs) and one FDES() call sharded inside one process (FDES_B200_GPUS).
N=${1:-2}
mkdir -p gpurun_out/multi
nvidia-smi -L | head -$N
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/multi/bench_${N}gpu.json 2> gpurun_out/multi/bench_${N}gpu.err
tail -c 400 gpurun_out/multi/bench_${N}gpu.err
python - "$N" <<'PY' 2>/dev/null | tee gpurun_out/multi/fdes_env_gpus_$N.txt
import numpy as np, time, tempfile, os, pathlib, sys
sys.path.insert(0, ".")
import fdes_b200 as fb
from fdes_b200 import specimens
n = int(sys.argv[1])
tmp = pathlib.Path(tempfile.mkdtemp()); os.chdir(tmp)
cnf = tmp / "au.cnf"
atoms = np.ascontiguousarray(specimens.config_au_2048(cnf, frozen_phonons=32), np.float32)
ref = None
for gpus in (str(n), "1"):
    os.environ["FDES_B200_GPUS"] = gpus
    img = np.zeros((1, 1024, 1024), np.float32)
    best = 1e9
    for i in range(4):
        t0 = time.perf_counter()
        fb.cuda_FDES(0, 0, str(cnf), str(tmp / "m.bin"), str(tmp / "r.emd"), atoms, len(atoms), img)
        best = min(best, (time.perf_counter() - t0) * 1e3)
    if ref is None:
        ref = img.copy()
    rel = float(np.linalg.norm(img - ref) / np.linalg.norm(ref))
    print(f"FDES() one process, FDES_B200_GPUS={gpus}: Au 2048^2 x 32 configurations in {best:.1f} ms (best of 4), image vs the {n}-GPU run rel-L2 {rel:.2e}")
PY
