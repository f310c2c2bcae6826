#!/usr/bin/env python3
"""Developer report (GPU box): library vs numpy oracle vs reference goldens for tests/data/*.cnf.
Prints relative L2 errors; the pass/fail versions of these checks live in tests/."""
import argparse
import pathlib
import sys
import traceback

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
import fdes_b200 as fb  # noqa: E402
import fdes_oracle as orc  # noqa: E402


def rel(a, b):
    a = np.asarray(a).astype(np.complex128).ravel()
    b = np.asarray(b).astype(np.complex128).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--golden", default=str(ROOT / "tests" / "golden"))
    ap.add_argument("cases", nargs="*")
    a = ap.parse_args()
    cases = a.cases or sorted(p.stem for p in (ROOT / "tests" / "data").glob("*.cnf"))
    for name in cases:
        cnf = ROOT / "tests" / "data" / f"{name}.cnf"
        try:
            p, Z, xyz, dwf, occ = orc.read_cnf(str(cnf))
            gfile = pathlib.Path(a.golden) / f"{name}.npz"
            g = np.load(gfile) if gfile.exists() else None
            with fb.Simulation(cnf, want_exitwave=True) as sim:
                line = [f"{name:11s} m={sim.m1} m3={sim.m3} nAt={sim.nAt} nZ={sim.nZ}"]
                count = max(1, p.frPh)
                coords = [sim.jitter_next(0) for _ in range(count)]
                if g is not None:
                    line.append(f"xyz|ref {rel(np.stack(coords), g['xyz_cfg']):.2e}")
                res = orc.build_measurements(p, Z, xyz, dwf, occ)
                ps = res.params
                xyz0 = coords[0]
                bins = sim.bin_atoms(xyz0)
                ob = orc.bin_atoms(xyz0, ps)
                line.append("bins " + ("OK" if _bins_equal(bins, ob, Z, ps) else "MISMATCH"))
                V0 = sim.phase_grating(xyz0, 0)
                Zl = orc.list_of_elements(Z)
                V0o = orc.phase_grating(0, Z, Zl, xyz0, occ, ps.imPot, ps)
                line.append(f"V0|orc {rel(V0, V0o):.2e}")
                if g is not None:
                    line.append(f"V0|ref {rel(V0, g['V'][0]):.2e} (orc|ref {rel(V0o, g['V'][0]):.2e})")
                pe = sim.exit_wave(xyz0, 0)
                if g is not None:
                    line.append(f"psi|ref {rel(pe, g['psi_exit'][0]):.2e}")
            with fb.Simulation(cnf, want_exitwave=True) as sim:
                img, ew = sim.simulate()
                line.append(f"img|orc {rel(img, res.image):.2e} ew|orc {rel(ew, res.exitwave):.2e}")
                if g is not None:
                    line.append(f"img|ref {rel(img, g['image']):.2e} ew|ref {rel(ew, g['exitwave']):.2e} "
                                f"(orc|ref img {rel(res.image, g['image']):.2e} ew {rel(res.exitwave, g['exitwave']):.2e})")
            print("  ".join(line), flush=True)
        except Exception:
            print(f"{name}: FAILED")
            traceback.print_exc()


def _bins_equal(bins, ob, Z, ps):
    i1, i2, i3, _, _, ok = ob
    ok = ok & (i3 >= 0) & (i3 < ps.m3)
    got_ok = bins[:, 0] >= 0
    if not np.array_equal(got_ok, ok):
        return False
    return (np.array_equal(bins[ok, 0], i1[ok]) and np.array_equal(bins[ok, 1], i2[ok])
            and np.array_equal(bins[ok, 2], i3[ok]))


if __name__ == "__main__":
    main()
