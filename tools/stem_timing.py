#!/usr/bin/env python3
"""Developer script: one short STEM scan (for ncu launch lists of the STEM kernels)."""
import pathlib, sys, tempfile
import numpy as np
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import fdes_b200 as fb
from fdes_b200 import specimens
tmp = pathlib.Path(tempfile.mkdtemp())
atoms = specimens.config_srtio3_stem_512(tmp / "s.cnf")
pos = specimens.stem_raster(256)[:int(sys.argv[1]) if len(sys.argv) > 1 else 256]
det = np.array([[70.0, 200.0], [11.0, 22.0]], np.float32)
with fb.Simulation(tmp / "s.cnf", batch=32) as sim:
    sim.stem_scan(pos[:64], det)
    sig, ms = sim.stem_scan(pos, det)
print(len(pos), "probes", ms, "ms", len(pos) / ms * 1e3, "probes/s", sig.mean(0))
