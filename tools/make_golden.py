#!/usr/bin/env python3
"""TEST INFRASTRUCTURE -- generate the golden vectors under tests/golden/ by running the
UNMODIFIED reference (oracle/_ref/ref_harness, built from /root/reference by oracle/Makefile)
on a GPU box:

    gpurun -- 'python tools/make_golden.py --out gpurun_out/golden'
    cp gpurun_out/golden/*.npz tests/golden/

FDES ships no golden vectors or tests of its own, so these outputs of the
reference itself are the pins for both the numpy oracle (oracle/fdes_oracle.py) and the CUDA
product.  For every tests/data/<case>.cnf two reference runs are recorded:

  * `ref_harness run <cnf> <dir> 2`   -- stock buildMeasurements (src/crystalMaker.cu:227-424):
        image [n3][n2][n1], coherent exit-wave average [n3][m2][m1]
  * `ref_harness trace <cnf> <dir> 2` -- the k = 0 driver loop replayed with the reference's own
        functions: incident wave, jittered coordinates per phonon configuration, V and psi of the
        first two slices, exit wave per configuration, intensity before the detector tail, J.

The reference deposits atoms with float atomicAdd (src/crystalMaker.cu:100-119), so its output
is reproducible only to ~1e-7 relative; parity tolerances are those of BASELINE.md section 5.
"""
import argparse
import pathlib
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
HARNESS = ROOT / "oracle" / "_ref" / "ref_harness"


def read_meta(path):
    meta = {}
    for line in open(path):
        k, v = line.split()
        meta[k] = float(v)
    return meta


def run_case(cnf: pathlib.Path, out_dir: pathlib.Path, max_dump: int = 2):
    with tempfile.TemporaryDirectory() as td:
        td = pathlib.Path(td)
        local = td / cnf.name
        shutil.copy(cnf, local)
        if cnf.suffix == ".qsc":       # the QSTEM reader opens the .cfg unit cell relative to the cwd
            for cfg in cnf.parent.glob("*.cfg"):
                shutil.copy(cfg, td / cfg.name)
        d_run, d_tr = td / "run", td / "trace"
        subprocess.run([str(HARNESS), "run", str(local), str(d_run), "2"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        subprocess.run([str(HARNESS), "trace", str(local), str(d_tr), str(max_dump)], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        meta = read_meta(d_tr / "meta.txt")
        n1, n2, n3 = int(meta["n1"]), int(meta["n2"]), int(meta["n3"])
        m1, m2, m3 = int(meta["m1"]), int(meta["m2"]), int(meta["m3"])
        nAt, count = int(meta["nAt"]), max(1, int(meta["frPh"]))
        c = lambda f: np.fromfile(f, np.complex64).reshape(m2, m1)
        g = {"meta_keys": np.array(sorted(meta)), "meta_vals": np.array([meta[k] for k in sorted(meta)], np.float64)}
        g["image"] = np.fromfile(d_run / "image.f32", np.float32).reshape(n3, n2, n1)
        g["exitwave"] = np.fromfile(d_run / "exitwave.f32", np.float32).view(np.complex64).reshape(n3, m2, m1)
        g["psi_in"] = c(d_tr / "psi_in.c64")
        g["xyz_cfg"] = np.stack([np.fromfile(d_tr / f"xyz_cfg{j:03d}.f32", np.float32).reshape(nAt, 3)
                                 for j in range(count)])
        ns = min(max_dump, m3)
        g["V"] = np.stack([c(d_tr / f"V_s{s:04d}.c64") for s in range(ns)])
        g["psi_s"] = np.stack([c(d_tr / f"psi_s{s:04d}.c64") for s in range(ns)])
        g["psi_exit"] = np.stack([c(d_tr / f"psi_exit_cfg{j:03d}.c64") for j in range(count)])
        g["exitwave_avg_k0"] = c(d_tr / "exitwave_avg.c64")
        g["I_k0"] = np.ascontiguousarray(c(d_tr / "I_d.c64").real)
        g["J_k0"] = np.fromfile(d_tr / "J.f32", np.float32)[: n1 * n2].reshape(n2, n1)
        out_dir.mkdir(parents=True, exist_ok=True)
        if cnf.suffix == ".qsc":       # side-effect file of readQsc (src/rwQsc.cu:1084): what the reader understood
            shutil.copy(td / "ParamsUsedQsc.txt", out_dir / f"{cnf.stem}.ParamsUsedQsc.txt")
        np.savez_compressed(out_dir / f"{cnf.stem}.npz", **g)
        return meta


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "tests" / "golden"))
    ap.add_argument("cases", nargs="*")
    a = ap.parse_args()
    if not HARNESS.exists():
        sys.exit(f"{HARNESS} missing: run `make -C oracle` where /root/reference is mounted")
    data = ROOT / "tests" / "data"
    cases = [data / (c if "." in c else f"{c}.cnf") for c in a.cases] or sorted(data.glob("*.cnf")) + sorted(data.glob("*.qsc"))
    for cnf in cases:
        meta = run_case(cnf, pathlib.Path(a.out))
        print(f"golden {cnf.stem}: m={int(meta['m1'])} m3={int(meta['m3'])} nAt={int(meta['nAt'])} mode={int(meta['mode'])}")


if __name__ == "__main__":
    main()
