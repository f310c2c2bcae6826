"""TEST INFRASTRUCTURE -- CPU restatement (numpy / scipy.fft, float32) of the FDES forward
multislice hot path.  It is the checker for the CUDA product in ``fdes_b200/``; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it.
The product never does.

Parity status: PINNED against the reference itself.  FDES ships no golden vectors or tests
(SURVEY.md section 4), so the pins are outputs of the unmodified reference built by
``oracle/Makefile`` (``oracle/_ref/ref_harness``) and run on a B200; the vectors live in
``tests/golden/`` together with the generating script ``tools/make_golden.sh``.

Every function cites the reference file:line (relative to /root/reference) it restates.
All arithmetic is float32 with the reference's operation order; FFTs are unnormalised with
the forward sign exp(-2 pi i ...), identical to cuFFT C2C on an array indexed [i2][i1].
"""
from __future__ import annotations

import dataclasses
import math
import os
import pathlib
import re
from typing import List, Optional, Tuple

import numpy as np
import scipy.fft as sfft

f32 = np.float32
c64 = np.complex64
FLT_EPSILON = f32(1.1920929e-07)
FLT_MIN = f32(1.17549435e-38)
_HERE = pathlib.Path(__file__).resolve().parent
_WORKERS = int(os.environ.get("FDES_ORACLE_WORKERS", os.cpu_count() or 1))


# --------------------------------------------------------------------------------------
# parameters  (include/paramStructure.h:48-162)
# --------------------------------------------------------------------------------------
ABERRATION_KEYS = ["C1", "A1", "A2", "B2", "C3", "A3", "S3", "A4", "B4", "D4", "C5", "A5", "R5", "S5"]


@dataclasses.dataclass
class Params:
    # defaults: defaultParams, src/paramStructure.cu:490-586
    E0: f32 = f32(200e3)
    gamma: f32 = f32(1.3913902)
    lam: f32 = f32(2.507934e-12)
    sigma: f32 = f32(7288400.5)
    ab0: dict = dataclasses.field(default_factory=lambda: {k: f32(0) for k in ABERRATION_KEYS})
    ab1: dict = dataclasses.field(default_factory=lambda: {k: f32(0) for k in ABERRATION_KEYS})
    defocspread: f32 = f32(0)
    illangle: f32 = f32(0)
    mtfa: f32 = f32(1)
    mtfb: f32 = f32(0)
    mtfc: f32 = f32(0)
    mtfd: f32 = f32(0)
    ObjAp: f32 = f32(11.1e-3)
    mode: int = 0
    m1: int = 4
    m2: int = 4
    m3: int = 1
    d1: f32 = f32(0.25e-10)
    d2: f32 = f32(0.25e-10)
    d3: f32 = f32(2e-10)
    dn1: int = 1
    dn2: int = 1
    n1: int = 2
    n2: int = 2
    n3: int = 1000
    frPh: int = 0
    pD: f32 = f32(0)
    subSlTh: f32 = f32(2e-10)
    tiltspec: np.ndarray = None
    tiltbeam: np.ndarray = None
    defoci: np.ndarray = None
    tilt_off: Tuple[f32, f32, f32] = (f32(0), f32(0), f32(0))
    doBeamTilt: bool = False
    imPot: f32 = f32(0)
    pi: f32 = f32(3.141592654)  # allocParams, src/paramStructure.cu:700

    def __post_init__(self):
        self.ab0 = dict(self.ab0)
        self.ab0["C1"] = f32(-6.1334e-008)
        self.ab0["C3"] = f32(1e-3)
        if self.tiltspec is None:
            self.tiltspec = np.zeros(2 * 1000, f32)
            self.tiltbeam = np.zeros(2 * 1000, f32)
            self.defoci = np.zeros(1000, f32)

    def copy(self) -> "Params":
        q = dataclasses.replace(self)
        q.ab0, q.ab1 = dict(self.ab0), dict(self.ab1)
        q.tiltspec, q.tiltbeam, q.defoci = self.tiltspec.copy(), self.tiltbeam.copy(), self.defoci.copy()
        return q


def _sscanf_g(tok: str) -> Optional[f32]:
    """%g of sscanf: longest numeric prefix (strtof semantics), None if no conversion."""
    m = re.match(r"[-+]?(\d+\.?\d*([eE][-+]?\d+)?|\.\d+([eE][-+]?\d+)?|inf|nan)", tok, re.I)
    if not m:
        return None
    return f32(float(m.group(0)))


def _sscanf_i(tok: str) -> Optional[int]:
    """%i / %d of sscanf (decimal only is what the inputs use)."""
    m = re.match(r"[-+]?\d+", tok)
    return int(m.group(0)) if m else None


def read_cnf(path: str, atoms_from_external: bool = False):
    """readConfig + numberOfAtoms + readCoordinates + getParams
    (src/paramStructure.cu:42-302, 588-635, 1019-1077).
    Returns (Params after consitentParams, Z int32[nAt], xyz f32[nAt,3], DWF f32[nAt], occ f32[nAt]).
    Lines are read with a 100-byte fgets like the reference (:46,66): longer lines continue
    as a new 'line'."""
    p = Params()
    raw = open(path, "rb").read().decode("latin-1")
    ts_i = tb_i = df_i = 0

    def chunks(text, size):
        for line in text.split("\n"):
            line = line + "\n"
            while len(line) > size - 1:
                yield line[: size - 1]
                line = line[size - 1 :]
            yield line

    last_field = ""
    lines = list(chunks(raw, 100))
    if raw.endswith("\n"):
        lines = lines[:-1]
    for line in lines:
        toks = line.split()
        field = toks[0] if toks else last_field  # sscanf leaves fieldName untouched on blank lines
        last_field = field
        if not toks:
            # NOTE: the reference re-processes the previous fieldName with the *new* (blank) line;
            # every sscanf then fails to convert, so nothing changes -- except the array counters.
            if field.startswith("specimen_tilt:"):
                ts_i += 1
            if field.startswith("beam_tilt:"):
                tb_i += 1
            if field.startswith("defoci:"):
                df_i += 1
            continue
        args = toks[1:]

        def g(i):
            return _sscanf_g(args[i]) if len(args) > i else None

        def setf(name, v):
            if v is not None:
                setattr(p, name, v)

        def seti(name, i=0):
            v = _sscanf_i(args[i]) if len(args) > i else None
            if v is not None:
                setattr(p, name, v)

        if field.startswith("voltage:"):
            setf("E0", g(0))
        for key in ABERRATION_KEYS:
            if field.startswith(key + ":"):
                v0 = g(0)
                if v0 is not None:
                    p.ab0[key] = v0
                    if key not in ("C1", "C3", "C5"):
                        v1 = g(1)
                        if v1 is not None:
                            p.ab1[key] = v1
        if field.startswith("focus_spread"):  # strncmp(...,12)
            setf("defocspread", g(0))
        if field.startswith("illumination_angle:"):
            setf("illangle", g(0))
        if field.startswith("mtf_a:"):
            setf("mtfa", g(0))
        if field.startswith("mtf_b:"):
            setf("mtfb", g(0))
        if field.startswith("mtf_c:"):
            setf("mtfc", g(0))
        if field.startswith("mtf_d:"):
            setf("mtfd", g(0))
        if field.startswith("objective_aperture:"):
            setf("ObjAp", g(0))
        if field.startswith("sample_size_x:"):
            seti("m1")
        if field.startswith("sample_size_y:"):
            seti("m2")
        if field.startswith("sample_size_z:"):
            seti("m3")
        if field.startswith("pixel_size_x:"):
            setf("d1", g(0))
        if field.startswith("pixel_size_y:"):
            setf("d2", g(0))
        if field.startswith("pixel_size_z:"):
            setf("d3", g(0))
        if field.startswith("border_size_x:"):
            seti("dn1")
        if field.startswith("border_size_y:"):
            seti("dn2")
        if field.startswith("image_size_x:"):
            seti("n1")
        if field.startswith("image_size_y:"):
            seti("n2")
        if field.startswith("image_size_z:"):
            seti("n3")
        if field.startswith("specimen_tilt:"):
            a, b = g(0), g(1)
            if a is not None:
                p.tiltspec[2 * ts_i] = a
                if b is not None:
                    p.tiltspec[2 * ts_i + 1] = b
            ts_i += 1
        if field.startswith("beam_tilt:"):
            a, b = g(0), g(1)
            if a is not None:
                p.tiltbeam[2 * tb_i] = a
                if b is not None:
                    p.tiltbeam[2 * tb_i + 1] = b
            tb_i += 1
        if field.startswith("defoci:"):
            a = g(0)
            if a is not None:
                p.defoci[df_i] = a
            df_i += 1
        if field.startswith("absorptive_potential_factor:"):
            setf("imPot", g(0))
        if field.startswith("pixel_dose:"):
            setf("pD", g(0))
        if field.startswith("frozen_phonons:"):
            seti("frPh")
        if field.startswith("subpixel_size_z:"):
            setf("subSlTh", g(0))
        if field.startswith("specimen_tilt_offset_x:"):
            v = g(0)
            if v is not None:
                p.tilt_off = (v, p.tilt_off[1], p.tilt_off[2])
        if field.startswith("specimen_tilt_offset_y:"):
            v = g(0)
            if v is not None:
                p.tilt_off = (p.tilt_off[0], v, p.tilt_off[2])
        if field.startswith("specimen_tilt_offset_z:"):
            v = g(0)
            if v is not None:
                p.tilt_off = (p.tilt_off[0], p.tilt_off[1], v)
        if field.startswith("mode:"):
            seti("mode")

    # atoms: numberOfAtoms/readCoordinates use a 200-byte fgets (src/paramStructure.cu:1019-1077).
    # The field name is only refreshed when sscanf("%s") converts and the loop runs until feof():
    # a file ending in "atom: ...\n" therefore goes round once more with a failed fgets, the
    # stale field name "atom:" and the stale line buffer ("#tom: ..." after resetLine, same
    # numbers) -> the last atom is read twice.  A blank line after an atom line also re-dispatches
    # "atom:" but converts nothing (the reference then keeps uninitialised memory; zeros here).
    Z, xyz, dwf, occ = [], [], [], []
    if not atoms_from_external:
        alines = list(chunks(raw, 200))
        if raw.endswith("\n"):
            alines[-1] = None      # the failed fgets at end-of-file
        else:
            alines[-1] = alines[-1][:-1]
        field, stale = "", ""
        for line in alines:
            if line is None:
                line = stale
            else:
                toks = line.split()
                if toks:
                    field = toks[0]
            if field.startswith("atom:"):
                vals = line.split()[1:7]
                zi = _sscanf_i(vals[0]) if vals else None
                nums = [(_sscanf_g(v) if zi is not None else None) for v in vals[1:6]]
                nums += [None] * (5 - len(nums))
                for i in range(5):          # sscanf stops at the first failed conversion
                    if nums[i] is None:
                        nums[i:] = [None] * (5 - i)
                        break
                nums = [f32(0) if v is None else v for v in nums]
                Z.append(zi if zi is not None else 0)
                xyz.append(nums[0:3])
                dwf.append(nums[3])
                occ.append(nums[4])
            stale = "#" + line[1:]
    p2 = p.copy()
    n3 = p2.n3
    p2.tiltspec, p2.tiltbeam, p2.defoci = p.tiltspec[: 2 * n3].copy(), p.tiltbeam[: 2 * n3].copy(), p.defoci[:n3].copy()
    consistent_params(p2)
    return (p2, np.asarray(Z, np.int32), np.asarray(xyz, f32).reshape(-1, 3), np.asarray(dwf, f32), np.asarray(occ, f32))


def consistent_params(p: Params) -> None:
    """consitentParams, src/paramStructure.cu:637-673 (float32, same operation order;
    sqrt is the double overload applied to a float expression, then narrowed)."""
    E0 = f32(p.E0)
    m0, c, e, h = f32(9.1093822), f32(2.9979246), f32(1.6021766), f32(6.6260696)
    pi = f32(p.pi)
    p.gamma = f32(f32(1) + f32(f32(f32(f32(E0 * e) / m0) / c) / c) * f32(1e-4))
    s1 = f32(math.sqrt(float(f32(f32(f32(2) * m0) * e))))
    inner = f32(f32(1) + f32(f32(f32(f32(f32(E0 * e) / f32(2)) / m0) / c) / c) * f32(1e-4))
    s2 = f32(math.sqrt(float(f32(E0 * inner))))
    # h / sqrt(..) * 1e-9f / sqrt(..): sqrt(float) in CUDA host code resolves to the float overload
    p.lam = f32(f32(f32(h / s1) * f32(1e-9)) / s2)
    p.sigma = f32(f32(f32(f32(f32(f32(f32(f32(2) * pi) * p.gamma) * p.lam) * m0) * e) / h) / h) * f32(1e18)
    p.sigma = f32(p.sigma)
    p.m1 = p.n1 + 2 * p.dn1
    p.m2 = p.n2 + 2 * p.dn2
    flag = f32(0)
    for j in range(2 * p.n3):
        flag = f32(flag + abs(p.tiltbeam[j]))
    p.doBeamTilt = not (flag < f32(FLT_MIN * f32(p.n3 * 2)))


def sub_slice_ratio(slice_: f32, sub: f32) -> f32:
    """subSliceRatio, src/crystalMaker.cu:720-728."""
    ratio = f32(1)
    if sub > f32(1e-12) and sub < slice_:
        ratio = f32(math.ceil(float(f32(slice_ / sub))))
    return ratio


def set_sub_slices(p: Params, ratio: f32) -> None:
    """setSubSlices, src/crystalMaker.cu:730-743."""
    p.m3 = int(f32(f32(p.m3) * ratio))
    p.d3 = f32(p.d3 / ratio)


# --------------------------------------------------------------------------------------
# index helpers  (include/coordArithmetic.h:28-40)
# --------------------------------------------------------------------------------------
def iw(m: int) -> np.ndarray:
    i = np.arange(m, dtype=np.int32)
    return np.where(i > m // 2, i - m, i).astype(np.int32)


def ow(m: int) -> np.ndarray:
    return (np.arange(m, dtype=np.int32) - m // 2).astype(np.int32)


def roundf(x: np.ndarray) -> np.ndarray:
    """C roundf: half away from zero, exact (x - trunc(x) is exact in binary fp)."""
    x = np.asarray(x, f32)
    t = np.trunc(x)
    return (t + np.sign(x) * (np.abs(x - t) >= f32(0.5))).astype(f32)


def fft2(a: np.ndarray) -> np.ndarray:
    return sfft.fft2(a.astype(c64, copy=False), workers=_WORKERS).astype(c64, copy=False)


def ifft2(a: np.ndarray) -> np.ndarray:
    """unnormalised inverse (CUFFT_INVERSE)."""
    return sfft.ifft2(a.astype(c64, copy=False), norm="forward", workers=_WORKERS).astype(c64, copy=False)


# --------------------------------------------------------------------------------------
# Kirkland table and species list
# --------------------------------------------------------------------------------------
_KIRK = None


def kirkland_table() -> np.ndarray:
    """103 x 12 float32 (a0 b0 a1 b1 a2 b2 c0 d0 c1 d1 c2 d2), src/projectedPotential.cu:94-2983."""
    global _KIRK
    if _KIRK is None:
        _KIRK = np.loadtxt(_HERE / "kirkland_table.txt", dtype=np.float64).astype(f32)
        assert _KIRK.shape == (103, 12)
    return _KIRK


def kirkland_params(Z: int):
    """parametersKirkland_d incl. the unknown-Z fallback a=0,b=1,c=1,d=0
    (src/projectedPotential.cu:2984-3009)."""
    if 1 <= Z <= 103:
        r = kirkland_table()[Z - 1]
        return r[[0, 2, 4]], r[[1, 3, 5]], r[[6, 8, 10]], r[[7, 9, 11]]
    return np.zeros(3, f32), np.ones(3, f32), np.ones(3, f32), np.zeros(3, f32)


def list_of_elements(Z: np.ndarray) -> List[int]:
    """listOfElements, src/crystalMaker.cu:539-570: unique Z in first-appearance order."""
    out: List[int] = []
    for z in Z.tolist():
        if z not in out:
            out.append(z)
    return out


# --------------------------------------------------------------------------------------
# atoms -> grid  (squareAtoms_d, src/crystalMaker.cu:73-134)
# --------------------------------------------------------------------------------------
def bin_atoms(xyz: np.ndarray, p: Params):
    """Per-atom (i1, i2, i3, r1, r2, ok_xy).  Bit-exact integer targets of the product."""
    xyz = np.asarray(xyz, f32)
    m1, m2, m3 = p.m1, p.m2, p.m3
    x1 = f32(xyz[:, 0] / f32(p.d1)) + f32(f32(m1) * f32(0.5))
    x1 = (x1.astype(f32) - f32(0.5)).astype(f32)
    x2 = f32(xyz[:, 1] / f32(p.d2)) + f32(f32(m2) * f32(0.5))
    x2 = (x2.astype(f32) - f32(0.5)).astype(f32)
    x3 = f32(xyz[:, 2] / f32(p.d3)) + f32(f32(m3) * f32(0.5))
    x3 = (x3.astype(f32) - f32(0.5)).astype(f32)
    i3 = roundf(x3).astype(np.int64).clip(-(2**31), 2**31 - 1).astype(np.int32)
    ok = (x1 > f32(1)) & (x1 < f32(m1 - 2)) & (x2 > f32(1)) & (x2 < f32(m2 - 2))
    i1 = roundf(x1).astype(np.int64).clip(-(2**31), 2**31 - 1).astype(np.int32)
    i2 = roundf(x2).astype(np.int64).clip(-(2**31), 2**31 - 1).astype(np.int32)
    r1 = (x1 - i1.astype(f32)).astype(f32)
    r2 = (x2 - i2.astype(f32)).astype(f32)
    return i1, i2, i3, r1, r2, ok


def square_atoms(Zarr, Z0, xyz, occ, imPot, s, p: Params, bins=None) -> np.ndarray:
    """Density of species Z0 in slice s, complex64 [m2, m1] (bilinear deposition)."""
    m1, m2 = p.m1, p.m2
    i1, i2, i3, r1, r2, ok = bins if bins is not None else bin_atoms(xyz, p)
    sel = ok & (Zarr == Z0) & (i3 == s)
    V = np.zeros(m1 * m2, f32)
    if sel.any():
        i1, i2, r1, r2, oc = i1[sel], i2[sel], r1[sel], r2[sel], np.asarray(occ, f32)[sel]
        g1 = np.where(r1 < 0, -1, 1).astype(np.int32)
        g2 = np.where(r2 < 0, -1, 1).astype(np.int32)
        a1, a2 = np.abs(r1), np.abs(r2)
        one = f32(1)
        w = [((one - a1) * (one - a2)).astype(f32) * oc, ((one - a1) * a2).astype(f32) * oc,
             (a1 * a2).astype(f32) * oc, (a1 * (one - a2)).astype(f32) * oc]
        idx = [i2 * m1 + i1, (i2 + g2) * m1 + i1, (i2 + g2) * m1 + i1 + g1, i2 * m1 + i1 + g1]
        for ww, jj in zip(w, idx):
            np.add.at(V, jj, ww.astype(f32))
    re = V.reshape(m2, m1)
    return (re + 1j * (re * f32(imPot))).astype(c64)


def scattering_factor(Z: int, p: Params) -> np.ndarray:
    """projectedPotential_d then divideBySinc: real float32 [m2, m1]
    (src/projectedPotential.cu:30-73, src/crystalMaker.cu:136-158)."""
    m1, m2 = p.m1, p.m2
    a, b, c, d = kirkland_params(Z)
    d1 = f32(f32(1e10) * f32(p.d1))
    d2 = f32(f32(1e10) * f32(p.d2))
    i1 = iw(m1).astype(f32)[None, :]
    i2 = iw(m2).astype(f32)[:, None]
    q1 = (i1 / f32(d1 * f32(m1))).astype(f32)
    q2 = (i2 / f32(d2 * f32(m2))).astype(f32)
    qsq = (q1 * q1 + q2 * q2).astype(f32)
    Vz = np.zeros_like(qsq)
    for k in range(3):
        Vz = (Vz + (a[k] / (qsq + b[k]).astype(f32) + c[k] * np.exp((-d[k] * qsq).astype(f32)).astype(f32)).astype(f32)).astype(f32)
    num = f32(f32(4.78776452e-9) * f32(p.sigma))
    den = f32(f32(d1 * d2) * f32(m1 * m2))
    V = ((Vz * num).astype(f32) / den).astype(f32)
    pi = f32(p.pi)
    x = ((i1 / f32(m1)).astype(f32) * pi).astype(f32)
    x = ((x + FLT_EPSILON) / (np.sin(x).astype(f32) + FLT_EPSILON)).astype(f32)
    y = (pi * (i2 / f32(m2)).astype(f32)).astype(f32)
    x = (x * ((y + FLT_EPSILON) / (np.sin(y).astype(f32) + FLT_EPSILON)).astype(f32)).astype(f32)
    return (V * x).astype(f32)


def phase_grating(s, Zarr, Zlist, xyz, occ, imPot, p: Params, bins=None, sf_cache=None) -> np.ndarray:
    """phaseGrating, src/crystalMaker.cu:507-536.  Note the reference plan is
    cufftPlan2d(m1, m2) (:575) -- identical to [m2, m1] only for square grids."""
    assert p.m1 == p.m2, "reference phaseGrating transposes the plan for non-square grids"
    V = np.zeros((p.m2, p.m1), c64)
    for Z0 in Zlist:
        rho = square_atoms(Zarr, Z0, xyz, occ, imPot, s, p, bins)
        if sf_cache is not None:
            if Z0 not in sf_cache:
                sf_cache[Z0] = scattering_factor(Z0, p)
            G = sf_cache[Z0]
        else:
            G = scattering_factor(Z0, p)
        if not rho.any():
            continue  # FFT of exact zeros is exact zeros in the reference as well
        V = (V + ifft2(fft2(rho) * G)).astype(c64)
    return V


# --------------------------------------------------------------------------------------
# transmit + propagate  (src/multisliceSimulation.cu:41-52, 225-274, 538-611)
# --------------------------------------------------------------------------------------
def band_mask(p: Params) -> np.ndarray:
    """zeroHighFreq keep-mask, src/multisliceSimulation.cu:225-250."""
    mind = f32(min(p.m1, p.m2))
    i1 = iw(p.m1).astype(np.int64)[None, :]
    i2 = iw(p.m2).astype(np.int64)[:, None]
    v = ((i1 * i1 + i2 * i2).astype(f32) * f32(9) / f32(mind * mind)).astype(f32)
    return ~(v > f32(1))


def bandwidth_limit(f: np.ndarray, p: Params, mask=None) -> np.ndarray:
    """bandwidthLimit, src/multisliceSimulation.cu:552-560."""
    mask = band_mask(p) if mask is None else mask
    F = fft2(f) * mask
    return (ifft2(F) * f32(f32(1) / f32(p.m1 * p.m2))).astype(c64)


def fresnel_propagator(p: Params, mask=None) -> np.ndarray:
    """fresnelPropagatorDevice + zeroHighFreq + Csscal(1/N),
    src/multisliceSimulation.cu:253-274, 594-603."""
    m1, m2 = p.m1, p.m2
    d3 = f32(p.d3)
    i1 = iw(m1).astype(f32)[None, :]
    i2 = iw(m2).astype(f32)[:, None]
    t1 = ((i1 / f32(m1)).astype(f32) * f32(d3 / f32(p.d1))).astype(f32)
    t2 = ((i2 / f32(m2)).astype(f32) * f32(d3 / f32(p.d2))).astype(f32)
    ld = f32(f32(p.lam) / d3)
    ph = ((f32(-p.pi) * (t1 * t1 + t2 * t2).astype(f32)).astype(f32) * ld).astype(f32)
    P = (np.cos(ph).astype(f32) + 1j * np.sin(ph).astype(f32)).astype(c64)
    mask = band_mask(p) if mask is None else mask
    return (P * mask * f32(f32(1) / f32(m1 * m2))).astype(c64)


def potential2transmission(V: np.ndarray) -> np.ndarray:
    """potential2Transmission, src/multisliceSimulation.cu:41-52."""
    e = np.exp((-V.imag).astype(f32)).astype(f32)
    return ((e * np.cos(V.real.astype(f32))).astype(f32) + 1j * (e * np.sin(V.real.astype(f32))).astype(f32)).astype(c64)


def multiply_elementwise(f0: np.ndarray, f1: np.ndarray) -> np.ndarray:
    """multiplyElementwise (3-multiply form), src/complexMath.cu:44-62."""
    a, b = f0.real.astype(f32), f0.imag.astype(f32)
    c, d = f1.real.astype(f32), f1.imag.astype(f32)
    k = (a * (c + d)).astype(f32)
    dd = (d * (a + b)).astype(f32)
    cc = (c * (b - a)).astype(f32)
    return ((k - dd).astype(f32) + 1j * (k + cc).astype(f32)).astype(c64)


def forward_propagation(psi, V, p: Params, P=None, mask=None):
    """forwardPropagation, src/multisliceSimulation.cu:538-549.  Returns new psi."""
    mask = band_mask(p) if mask is None else mask
    P = fresnel_propagator(p, mask) if P is None else P
    t = potential2transmission(V)
    t = bandwidth_limit(t, p, mask)
    t = multiply_elementwise(t, psi)
    t = ifft2(multiply_elementwise(fft2(t), P))
    return t.astype(c64)


# --------------------------------------------------------------------------------------
# float64 evaluation of the same model (error budget of long slice chains)
# --------------------------------------------------------------------------------------
def exit_wave_fp64(p_in: Params, Z, xyz, occ, psi0: Optional[np.ndarray] = None) -> np.ndarray:
    """Exit wave (no frozen phonons, k = 0; plane wave, or the incident wave psi0 -- e.g. the float32
    probe of incoming_wave -- taken as exact) of the SAME model evaluated in float64 /
    complex128: identical float32 parameters (lambda, sigma, pixel sizes), identical bin decisions
    and bilinear weights (they are exact float32 quantities, src/crystalMaker.cu:85-119), but the
    scattering factors, exponentials and all transforms in double precision.  A float32 program that
    follows A.2-A.4 differs from this by its accumulated rounding only, which is what the parity
    tests of deep slice chains (400 sub-slices) compare the library and the reference against."""
    p = p_in.copy()
    set_sub_slices(p, sub_slice_ratio(p.d3, p.subSlTh))
    m1, m2 = p.m1, p.m2
    assert m1 == m2
    f64, c128 = np.float64, np.complex128
    Zl = list_of_elements(Z)
    i1, i2, i3, r1, r2, ok = bin_atoms(xyz, p)
    i1f = iw(m1).astype(f64)[None, :]
    i2f = iw(m2).astype(f64)[:, None]
    d1A, d2A = f64(f32(1e10) * f32(p.d1)), f64(f32(1e10) * f32(p.d2))
    qsq = (i1f / (d1A * m1)) ** 2 + (i2f / (d2A * m2)) ** 2
    eps = f64(FLT_EPSILON)
    pi = f64(f32(p.pi))
    sx = (pi * i1f / m1 + eps) / (np.sin(pi * i1f / m1) + eps)
    sy = (pi * i2f / m2 + eps) / (np.sin(pi * i2f / m2) + eps)
    G = {}
    for Z0 in Zl:
        a, b, c, d = [np.asarray(v, f64) for v in kirkland_params(Z0)]
        fz = sum(a[k] / (qsq + b[k]) + c[k] * np.exp(-d[k] * qsq) for k in range(3))
        G[Z0] = fz * (f64(f32(4.78776452e-9)) * f64(p.sigma)) / (d1A * d2A * f64(m1 * m2)) * sx * sy
    i1i = iw(m1).astype(np.int64)[None, :]
    i2i = iw(m2).astype(np.int64)[:, None]
    mask = ~(((i1i * i1i + i2i * i2i).astype(f32) * f32(9) / f32(f32(min(m1, m2)) * f32(min(m1, m2)))).astype(f32) > f32(1))
    d3 = f64(f32(p.d3))
    t1 = (i1f / m1) * (d3 / f64(f32(p.d1)))
    t2 = (i2f / m2) * (d3 / f64(f32(p.d2)))
    P = np.exp(-1j * pi * (t1 * t1 + t2 * t2) * (f64(f32(p.lam)) / d3)) * mask / f64(m1 * m2)
    invN = 1.0 / f64(m1 * m2)
    fft = lambda a: sfft.fft2(a, workers=_WORKERS)
    ifft = lambda a: sfft.ifft2(a, norm="forward", workers=_WORKERS)
    occ64 = np.asarray(occ, f32).astype(f64)
    psi = np.ones((m2, m1), c128) if psi0 is None else np.asarray(psi0).astype(c128)
    Zarr = np.asarray(Z)
    imPot = f64(f32(p.imPot))
    for s in range(p.m3):
        V = np.zeros((m2, m1), c128)
        for Z0 in Zl:
            sel = ok & (Zarr == Z0) & (i3 == s)
            if not sel.any():
                continue
            a1, a2 = np.abs(r1[sel]).astype(f64), np.abs(r2[sel]).astype(f64)
            g1 = np.where(r1[sel] < 0, -1, 1)
            g2 = np.where(r2[sel] < 0, -1, 1)
            j1, j2, oc = i1[sel].astype(np.int64), i2[sel].astype(np.int64), occ64[sel]
            rho = np.zeros(m1 * m2, f64)
            np.add.at(rho, j2 * m1 + j1, (1 - a1) * (1 - a2) * oc)
            np.add.at(rho, (j2 + g2) * m1 + j1, (1 - a1) * a2 * oc)
            np.add.at(rho, (j2 + g2) * m1 + j1 + g1, a1 * a2 * oc)
            np.add.at(rho, j2 * m1 + j1 + g1, a1 * (1 - a2) * oc)
            V += ifft(fft(rho.reshape(m2, m1) * (1 + 1j * imPot)) * G[Z0])
        t = np.exp(-V.imag) * np.exp(1j * V.real)
        t = ifft(fft(t) * mask) * invN
        psi = ifft(fft(t * psi) * P)
    return psi


# --------------------------------------------------------------------------------------
# incident wave, lens, detector  (src/multisliceSimulation.cu:89-156, 277-442, 563-622;
# src/crystalMaker.cu:187-224, 579-613, 700-718; src/complexMath.cu:510-557)
# --------------------------------------------------------------------------------------
def lens_function(p: Params, k: int):
    """CTF factor and aperture of multiplyLensFunction, src/multisliceSimulation.cu:277-343."""
    m1, m2 = p.m1, p.m2
    i1 = iw(m1).astype(f32)[None, :]
    i2 = (-iw(m2)).astype(f32)[:, None]
    nu1 = ((i1 / f32(m1)).astype(f32) * f32(f32(p.lam) / f32(p.d1))).astype(f32)
    nu2 = ((i2 / f32(m2)).astype(f32) * f32(f32(p.lam) / f32(p.d2))).astype(f32)
    nu1, nu2 = np.broadcast_arrays(nu1, nu2)
    phi = np.arctan2(nu2, nu1).astype(f32)
    nu = np.sqrt((nu1 * nu1 + nu2 * nu2).astype(f32)).astype(f32)
    a0, a1 = p.ab0, p.ab1

    def cs(n, key):
        return (a0[key] * np.cos((f32(n) * (phi - a1[key]).astype(f32)).astype(f32)).astype(f32)).astype(f32)

    def cs1(key):
        return (a0[key] * np.cos((phi - a1[key]).astype(f32)).astype(f32)).astype(f32)

    t5 = (f32(1.0 / 6.0) * (cs(6, "A5") + cs(4, "R5") + cs(2, "S5") + a0["C5"]).astype(f32)).astype(f32)
    t4 = (f32(0.2) * (cs(5, "A4") + cs1("B4") + cs(3, "D4")).astype(f32) + nu * t5).astype(f32)
    t3 = (f32(0.25) * (cs(4, "A3") + cs(2, "S3") + a0["C3"]).astype(f32) + nu * t4).astype(f32)
    t2 = (f32(1.0 / 3.0) * (cs(3, "A2") + cs1("B2")).astype(f32) + nu * t3).astype(f32)
    t1 = (f32(0.5) * (cs(2, "A1") + a0["C1"] + p.defoci[k]).astype(f32) + nu * t2).astype(f32)
    W = ((nu * nu).astype(f32) * t1).astype(f32)
    lam = f32(p.lam)
    damp = np.ones_like(nu)
    if p.mode == 0:
        dd = ((f32(p.defocspread) * nu).astype(f32) * nu / lam).astype(f32)
        damp = np.exp((f32(-2) * dd * dd).astype(f32)).astype(f32)
    arg = ((f32(2) * f32(p.pi)) * (W / lam).astype(f32)).astype(f32)
    re = (damp * np.cos(arg).astype(f32)).astype(f32)
    im = (damp * np.sin((-arg).astype(f32)).astype(f32)).astype(f32)
    ap = nu < f32(p.ObjAp)
    return (re + 1j * im).astype(c64), ap


def multiply_lens_function(psi, p: Params, k: int):
    ctf, ap = lens_function(p, k)
    return np.where(ap, psi * ctf, 0).astype(c64)


def fftshift2(a: np.ndarray) -> np.ndarray:
    """cufftShift2D_h, src/complexMath.cu:510-557: out[(i + m/2) mod m] = in[i] per axis
    (for i < m - m/2 move by +m/2, else by -(m - m/2))."""
    m2, m1 = a.shape
    return np.roll(np.roll(a, m1 // 2, axis=1), m2 // 2, axis=0)


def tilt_beam(psi, p: Params, k: int, flag: int):
    """tiltBeam_d, src/multisliceSimulation.cu:89-120."""
    i1 = ow(p.m1).astype(f32)[None, :]
    i2 = ow(p.m2).astype(f32)[:, None]
    x2 = f32(f32(p.lam) * f32(flag))
    x1 = ((i1 * f32(f32(p.d1) / x2)).astype(f32) * p.tiltbeam[2 * k + 1]).astype(f32)
    xx2 = ((i2 * f32(f32(p.d2) / x2)).astype(f32) * p.tiltbeam[2 * k]).astype(f32)
    ph = ((f32(2) * f32(p.pi)) * (x1 + xx2).astype(f32)).astype(f32)
    return (psi * (np.cos(ph).astype(f32) + 1j * np.sin(ph).astype(f32))).astype(c64)


def tapered_cosine_window(psi, p: Params):
    """taperedCosineWindow_d, src/multisliceSimulation.cu:123-156."""
    def win(m, dn):
        i = np.arange(m, dtype=f32)
        alpha = f32(f32(2) * f32(f32(dn) / f32(m)))
        x = (i / f32(m - 1)).astype(f32)
        w = np.ones(m, f32)
        lo = x < f32(alpha * f32(0.5))
        hi = (~lo) & (x > f32(f32(1) - f32(f32(0.5) * alpha)))
        with np.errstate(divide="ignore", invalid="ignore"):
            wlo = (f32(0.5) * (f32(1) + np.cos((f32(p.pi) * (f32(2) * x / alpha - f32(1)).astype(f32)).astype(f32)))).astype(f32)
            whi = (f32(0.5) * (f32(1) + np.cos((f32(p.pi) * (f32(2) * x / alpha + f32(1) - f32(2) / alpha).astype(f32)).astype(f32)))).astype(f32)
        w = np.where(lo, wlo, w)
        w = np.where(hi, whi, w)
        return w.astype(f32)
    return (psi * (win(p.m2, p.dn2)[:, None] * win(p.m1, p.dn1)[None, :]).astype(f32)).astype(c64)


def incoming_wave(p: Params, k: int, mask=None) -> np.ndarray:
    """incomingWave, src/multisliceSimulation.cu:563-591."""
    psi = np.ones((p.m2, p.m1), c64)
    if p.mode == 2:
        psi = multiply_lens_function(psi, p, k)
        psi = fftshift2(ifft2(psi))
        psi = bandwidth_limit(psi, p, mask)
        nrm = f32(np.sqrt(np.sum(np.abs(psi.astype(np.complex128)) ** 2)))
        alpha = f32(f32(np.sqrt(f32(p.n1 * p.n2))) / nrm)
        psi = (psi * alpha).astype(c64)
    if p.doBeamTilt:
        psi = tilt_beam(psi, p, k, 1)
        if p.mode in (0, 1):
            psi = tapered_cosine_window(psi, p)
            psi = bandwidth_limit(psi, p, mask)
    return psi


def apply_lens_function(psi, p: Params, k: int):
    """applyLensFunction, src/multisliceSimulation.cu:614-622."""
    out = ifft2(multiply_lens_function(fft2(psi), p, k))
    return (out * f32(f32(1) / f32(p.m1 * p.m2))).astype(c64)


def area_mask(p: Params) -> np.ndarray:
    """areaMask, src/multisliceSimulation.cu:468-510."""
    def w1d(m, dn):
        i = np.arange(m)
        w = np.ones(m, f32)
        with np.errstate(divide="ignore", invalid="ignore"):
            lo = (f32(0.5) * (f32(1) - np.cos((f32(3.1415927) * i.astype(f32) / f32(dn)).astype(f32)))).astype(f32)
            hi = (f32(0.5) * (f32(1) - np.cos((f32(3.1415927) * (m - i).astype(f32) / f32(dn)).astype(f32)))).astype(f32)
        w = np.where(i <= dn - 1, w * lo, w)
        w = np.where(i >= m - dn, w * hi, w)
        return w.astype(f32)
    return (w1d(p.m2, p.dn2)[:, None] * w1d(p.m1, p.dn1)[None, :]).astype(f32)


def diffraction_pattern(psi, p: Params, k: int, mask=None):
    """diffractionPattern, src/crystalMaker.cu:700-718 -> real intensity [m2, m1]."""
    if p.doBeamTilt:
        psi = tilt_beam(psi, p, k, -1)
    if p.mode == 1:
        am = area_mask(p)  # applyMaskFiltering, src/crystalMaker.cu:205-224 (blend towards 1)
        psi = ((f32(1) * (f32(1) - am)).astype(f32) + psi * am).astype(c64)
        psi = bandwidth_limit(psi, p, mask)
    F = fftshift2(fft2(psi))
    alpha = f32(np.sqrt(f32(f32(1) / f32(p.m1 * p.m2))))
    F = (F * alpha).astype(c64)
    return (F.real.astype(f32) ** 2 + F.imag.astype(f32) ** 2).astype(f32)


def mtf(p: Params) -> np.ndarray:
    """multiplyMtf factor, src/multisliceSimulation.cu:362-388."""
    nu1 = (iw(p.m1).astype(f32) / f32(p.m1)).astype(f32)[None, :]
    nu2 = (iw(p.m2).astype(f32) / f32(p.m2)).astype(f32)[:, None]
    r = np.sqrt((nu1 * nu1 + nu2 * nu2).astype(f32)).astype(f32)
    m = (f32(p.mtfa) * np.exp((f32(-p.mtfc) * r).astype(f32)).astype(f32)
         + f32(p.mtfb) * np.exp(((f32(-p.mtfd) * r).astype(f32) * r).astype(f32)).astype(f32)).astype(f32)
    a1 = (nu1 * f32(p.pi)).astype(f32)
    a2 = (nu2 * f32(p.pi)).astype(f32)
    s = (((np.sin(a1).astype(f32) + FLT_EPSILON) / (a1 + FLT_EPSILON)).astype(f32)
         * ((np.sin(a2).astype(f32) + FLT_EPSILON) / (a2 + FLT_EPSILON)).astype(f32)).astype(f32)
    return (m * s).astype(f32)


def spatial_incoherence(p: Params, k: int) -> np.ndarray:
    """multiplySpatialIncoherence / ...DP, src/multisliceSimulation.cu:391-442."""
    i1 = iw(p.m1).astype(f32)[None, :]
    i2 = iw(p.m2).astype(f32)[:, None]
    if p.mode == 0:
        lam = f32(p.lam)
        a = ((i1 / f32(p.m1)).astype(f32) * f32(lam / f32(p.d1))).astype(f32)
        b = ((i2 / f32(p.m2)).astype(f32) * f32(lam / f32(p.d2))).astype(f32)
        nusq = (a * a + b * b).astype(f32)
        damp = f32(f32(f32(p.pi) * f32(p.illangle)) * p.defoci[k])
        return np.exp(((-nusq) * damp * damp).astype(f32)).astype(f32)
    x1 = (i1 * f32(p.d1)).astype(f32)
    x2 = (i2 * f32(p.d2)).astype(f32)
    x1 = (x1 * x1 + x2 * x2).astype(f32)
    c = f32(f32(f32(p.pi) * f32(p.illangle)) / f32(p.lam))
    return np.exp((f32(-c) * c * x1).astype(f32)).astype(f32)


def add_noise_and_mtf(I: np.ndarray, p: Params, k: int, noise_normals: Optional[np.ndarray] = None):
    """addNoiseAndMtf + copyMiddleOut, src/crystalMaker.cu:579-613, src/optimFunctions.cu:109-121.
    I: complex64 [m2, m1] (imag 0).  Poisson/Anscombe noise needs the cuRAND stream; pass the
    per-pixel N(0,1) draws in noise_normals or keep pixel_dose == 0."""
    alpha = f32(f32(1) / f32(p.m1 * p.m2))
    F = fft2(I.astype(c64))
    if abs(f32(p.illangle)) > FLT_EPSILON:
        F = (F * spatial_incoherence(p, k)).astype(c64)
    if f32(p.pD) > FLT_EPSILON:
        assert noise_normals is not None, "pixel_dose > 0 needs the cuRAND normal draws"
        x = ifft2((F * alpha).astype(c64))
        fr = x.real.astype(f32).copy()
        dose = f32(p.pD)
        fi = (fr * dose).astype(f32)
        sel = fi > f32(1e-2)
        assert sel.all(), "restatement limited to images where every pixel draws (see build_measurements)"
        n = noise_normals.reshape(fi.shape).astype(f32)
        with np.errstate(invalid="ignore", divide="ignore"):
            v = (n * np.sqrt((f32(1) - np.exp((-fi / f32(0.777134)).astype(f32))).astype(f32))).astype(f32)
            v = (v + (f32(2) * np.sqrt((fi + f32(0.375)).astype(f32)) - f32(0.25) / np.sqrt(fi)).astype(f32)).astype(f32)
            v = roundf((f32(0.25) * v * v - f32(0.375)).astype(f32))
        v = np.where(v < FLT_MIN, f32(0), v)
        fr = np.where(sel, (v / dose).astype(f32), fr)
        F = fft2((fr + 1j * x.imag).astype(c64))
    F = (F * mtf(p)).astype(c64)
    out = ifft2((F * alpha).astype(c64))
    return out.real[p.dn2 : p.dn2 + p.n2, p.dn1 : p.dn1 + p.n1].astype(f32)


# --------------------------------------------------------------------------------------
# coordinates: tilt + frozen phonons  (src/crystalMaker.cu:28-48, 427-462)
# --------------------------------------------------------------------------------------
def tilt_coordinates(xyz: np.ndarray, t0, t1, t2) -> np.ndarray:
    """tiltCoordinates via cublasSrot semantics: x' = c x + s y ; y' = c y - s x with s = -sin."""
    xyz = np.array(xyz, f32, copy=True)

    def rot(a, b, ang):
        c, s = f32(np.cos(f32(ang))), f32(-np.sin(f32(ang)))
        x, y = xyz[:, a].copy(), xyz[:, b].copy()
        xyz[:, a] = (c * x + s * y).astype(f32)
        xyz[:, b] = (c * y - s * x).astype(f32)

    if abs(f32(t2)) > FLT_EPSILON:
        rot(0, 1, t2)
    if abs(f32(t1)) > FLT_EPSILON:
        rot(0, 2, t1)
    if abs(f32(t0)) > FLT_EPSILON:
        rot(1, 2, t0)
    return xyz


class Xorwow:
    """cuRAND XORWOW restated from the public algorithm (Marsaglia 2003 xorwow + Weyl 362437)
    with cuRAND's curand_init(seed, subsequence, 0) seeding: state scrambled from the seed, then
    skipped ahead by subsequence * 2^67 draws.  The skip is done by GF(2) matrix powers built
    here from the step function itself (no cuRAND tables).  Vectorised over subsequences.
    Third-party dependency: CUDA toolkit 12.9 curand_kernel.h (device API), call sites
    src/crystalMaker.cu:34,44,60."""

    def __init__(self, seed: int, nseq: int):
        s0 = (seed & 0xFFFFFFFF) ^ 0xAAD26B49
        s1 = ((seed >> 32) & 0xFFFFFFFF) ^ 0xF7DCEFDD
        t0 = (1099087573 * s0) & 0xFFFFFFFF
        t1 = (2591861531 * s1) & 0xFFFFFFFF
        d = (6615241 + t1 + t0) & 0xFFFFFFFF
        v = [(123456789 + t0) & 0xFFFFFFFF, 362436069 ^ t0, (521288629 + t1) & 0xFFFFFFFF, 88675123 ^ t1,
             (5783321 + t0) & 0xFFFFFFFF]
        self.v = np.tile(np.array(v, np.uint32)[:, None], (1, nseq))
        self.d = np.full(nseq, d, np.uint32)
        self._skip_sequences(np.arange(nseq, dtype=np.uint64))
        self.has_extra = np.zeros(nseq, bool)
        self.extra = np.zeros(nseq, f32)

    @staticmethod
    def _step_words(v):
        # v: list of 5 python ints (uint32) -> next state words (the xorshift part only)
        t = v[0] ^ (v[0] >> 2)
        n4 = (v[4] ^ ((v[4] << 4) & 0xFFFFFFFF)) ^ (t ^ ((t << 1) & 0xFFFFFFFF))
        return [v[1], v[2], v[3], v[4], n4 & 0xFFFFFFFF]

    @classmethod
    def _step_matrix(cls):
        # 160 x 160 GF(2) matrix as 160 column images (python ints of 160 bits)
        cols = []
        for b in range(160):
            v = [0] * 5
            v[b // 32] = 1 << (b % 32)
            w = cls._step_words(v)
            cols.append(sum(w[i] << (32 * i) for i in range(5)))
        return cols

    @staticmethod
    def _mat_mul(A, B):
        # (A o B): columns of result = A applied to columns of B
        out = []
        for col in B:
            acc, b = 0, 0
            while col:
                if col & 1:
                    acc ^= A[b]
                col >>= 1
                b += 1
            out.append(acc)
        return out

    def _skip_sequences(self, seq: np.ndarray):
        M = self._step_matrix()
        for _ in range(67):  # M^(2^67)
            M = self._mat_mul(M, M)
        maxbits = int(seq.max()).bit_length() if len(seq) and seq.max() > 0 else 0
        for bit in range(maxbits):
            sel = ((seq >> np.uint64(bit)) & np.uint64(1)).astype(bool)
            if sel.any():
                self._apply(M, sel)
            M = self._mat_mul(M, M)
        # Weyl counter: d += 362437 * 2^67 * n == 0 (mod 2^32): unchanged.

    def _apply(self, M, sel):
        cols = np.array([[(c >> (32 * i)) & 0xFFFFFFFF for i in range(5)] for c in M], np.uint32)  # [160,5]
        v = self.v[:, sel]
        out = np.zeros_like(v)
        for b in range(160):
            bitset = ((v[b // 32] >> np.uint32(b % 32)) & np.uint32(1)).astype(bool)
            if bitset.any():
                out[:, bitset] ^= cols[b][:, None]
        self.v[:, sel] = out

    def next_u32(self) -> np.ndarray:
        v = self.v
        t = v[0] ^ (v[0] >> np.uint32(2))
        n4 = (v[4] ^ (v[4] << np.uint32(4))) ^ (t ^ (t << np.uint32(1)))
        self.v = np.stack([v[1], v[2], v[3], v[4], n4])
        self.d = (self.d + np.uint32(362437)).astype(np.uint32)
        return (self.d + n4).astype(np.uint32)

    def normal(self) -> np.ndarray:
        """curand_normal: Box-Muller pairs, second value cached (curand_kernel.h / curand_normal.h).
        Device code uses logf / __sincosf-class intrinsics; float32 libm here agrees to ~1e-6."""
        out = np.empty(self.d.shape, f32)
        use = self.has_extra.copy()
        out[use] = self.extra[use]
        self.has_extra[use] = False
        need = ~use
        if need.any():
            # draw for all (cheap), keep where needed; advance only where needed
            v_save, d_save = self.v.copy(), self.d.copy()
            x = self.next_u32()
            y = self.next_u32()
            self.v[:, use] = v_save[:, use]
            self.d[use] = d_save[use]
            u = (x.astype(f32) * f32(2.3283064e-10) + f32(2.3283064e-10 / 2)).astype(f32)
            vv = (y.astype(f32) * f32(2.3283064e-10 * 6.2831855) + f32(2.3283064e-10 * 6.2831855 / 2)).astype(f32)
            s = np.sqrt((f32(-2) * np.log(u).astype(f32)).astype(f32)).astype(f32)
            out[need] = (s * np.sin(vv).astype(f32)).astype(f32)[need]
            self.extra[need] = (s * np.cos(vv).astype(f32)).astype(f32)[need]
            self.has_extra[need] = True
        return out


def atom_jitter(xyz: np.ndarray, dwf: np.ndarray, rng: Xorwow) -> np.ndarray:
    """atomJitter_d, src/crystalMaker.cu:37-48 (one generator per coordinate, index 3*atom+axis)."""
    n = rng.normal().reshape(-1, 3)
    sd = np.sqrt(np.asarray(dwf, f32)).astype(f32)[:, None]
    return (np.asarray(xyz, f32) + ((n * f32(0.112539540)).astype(f32) * sd).astype(f32)).astype(f32)


# --------------------------------------------------------------------------------------
# driver  (buildMeasurements, src/crystalMaker.cu:227-424)
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class Result:
    image: np.ndarray            # float32 [n3, n2, n1]
    exitwave: np.ndarray         # complex64 [n3, m2, m1]  (coherent phonon average)
    params: Params               # after sub-slicing
    I: Optional[np.ndarray] = None   # last k: intensity before addNoiseAndMtf, float32 [m2, m1]


def build_measurements(p_in: Params, Z, xyz, dwf, occ, *, jitter_coords: Optional[List[np.ndarray]] = None,
                       trace=None) -> Result:
    """Restates the k / j / s loops.  jitter_coords (one array per (k, j)) lets a test feed the
    GPU-generated phonon displacements; otherwise the numpy XORWOW restatement is used."""
    p = p_in.copy()
    ratio = sub_slice_ratio(p.d3, p.subSlTh)
    set_sub_slices(p, ratio)
    Z = np.asarray(Z, np.int32)
    occ = np.asarray(occ, f32)
    xyzTO = tilt_coordinates(xyz, *p.tilt_off)
    Zlist = list_of_elements(Z)
    count = p.frPh if p.frPh > 0 else 1
    rng = Xorwow(1, 3 * len(Z)) if (p.frPh > 0 and jitter_coords is None) else None
    mask = band_mask(p)
    P = fresnel_propagator(p, mask)
    sf_cache = {}
    alpha = f32(f32(1) / f32(count))
    image = np.zeros((p.n3, p.n2, p.n1), f32)
    exitwave = np.zeros((p.n3, p.m2, p.m1), c64)
    I = None
    # Poisson noise: one XORWOW stream per pixel, curand_init(1 + n3, pixel, 0) (src/crystalMaker.cu:295);
    # a stream advances only where a pixel draws (fi > 1e-2, :58) -- this restatement draws for every
    # pixel and therefore requires that condition to hold everywhere (checked in add_noise_and_mtf).
    noise_rng = Xorwow(1 + p.n3, p.m1 * p.m2) if f32(p.pD) > FLT_EPSILON else None
    for k in range(p.n3):
        I = np.zeros((p.m2, p.m1), f32)
        xyz_k = tilt_coordinates(xyzTO, p.tiltspec[2 * k], p.tiltspec[2 * k + 1], 0.0)
        for j in range(count):
            psi = incoming_wave(p, k, mask)
            if p.frPh > 0:
                xyzFP = jitter_coords[k * count + j] if jitter_coords is not None else atom_jitter(xyz_k, dwf, rng)
            else:
                xyzFP = xyz_k
            bins = bin_atoms(xyzFP, p)
            for s in range(p.m3):
                V = phase_grating(s, Z, Zlist, xyzFP, occ, p.imPot, p, bins, sf_cache)
                psi = forward_propagation(psi, V, p, P, mask)
                if trace is not None:
                    trace(k, j, s, V, psi)
            exitwave[k] = (exitwave[k] + alpha * psi).astype(c64)
            if p.mode == 0:
                q = apply_lens_function(psi, p, k)
                inten = (q.real.astype(f32) ** 2 + q.imag.astype(f32) ** 2).astype(f32)
            else:
                inten = diffraction_pattern(psi, p, k, mask)
            I = (I + alpha * inten).astype(f32)
        image[k] = add_noise_and_mtf(I.astype(c64), p, k, noise_rng.normal() if noise_rng is not None else None)
    return Result(image=image, exitwave=exitwave, params=p, I=I)
