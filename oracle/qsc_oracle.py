"""TEST INFRASTRUCTURE -- CPU restatement of the reference's .qsc / .cfg reader.

Only tests/ may import this (the product reader is fdes_b200/csrc/qsc.cpp).  It follows, in the
reference's own float32 / double mix:

  readparam                 qstem-libs/readparams.cpp:173-218   (position-based search, wraps once,
                                                                  strstr match, '%' comments)
  strnext                   qstem-libs/readparams.cpp:232-247
  readCFGCellParams         qstem-libs/fileio_fftw3.cpp:721-776
  readNextCFGAtom           qstem-libs/fileio_fftw3.cpp:908-975
  getZNumber                qstem-libs/fileio_fftw3.cpp:2299-2321
  readUnitCell (NCELL mode) qstem-libs/fileio_fftw3.cpp:1313-1657
  replicateUnitCell         qstem-libs/fileio_fftw3.cpp:1188-1305 (full occupancy, distinct sites)
  rotateVect                qstem-libs/matrixlib.cpp:599-635
  readQsc                   src/rwQsc.cu:8-1088, MULS -> params_t at :937-1001, atoms at :1003-1082

Pinned by tests/golden/qsc_*.txt: the `ParamsUsedQsc.txt` files the UNMODIFIED reference
(oracle/_ref/ref_harness) wrote for tests/data/*.qsc on a B200 box (tools/make_golden.py).
"""
import math
import os
import re

import numpy as np

from fdes_oracle import Params, consistent_params

f32 = np.float32
f64 = np.float64

EL_TABLE = ("H HeLiBeB C N O F NeNaMgAlSiP S Cl"
            "ArK CaScTiV CrMnFeCoNiCuZnGaGeAsSeBr"
            "KrRbSrY ZrNbMoTcRuRhPdAgCdInSnSbTe"
            "I XeCsBaLaCePrNdPmSmEuGdTbDyHoErTm"
            "YbLuHfTaW ReOsIrPtAuHgTlPbBiPoAtRn"
            "FrRaAcThPaU NpPuAmCmBkCfEsFmMdNoLr")

_NUM = re.compile(r"\s*([-+]?(\d+\.?\d*([eE][-+]?\d+)?|\.\d+([eE][-+]?\d+)?))")
_INT = re.compile(r"\s*([-+]?\d+)")


def _atof(s: str) -> float:
    m = _NUM.match(s)
    return float(m.group(1)) if m else 0.0


def _scan_g(s: str, default):
    m = _NUM.match(s)
    return f32(float(m.group(1))) if m else default


def _scan_d(s: str, default):
    m = _INT.match(s)
    return int(m.group(1)) if m else default


def _word(s: str) -> str:
    t = s.split()
    return t[0] if t else ""


class ParFile:
    """readparams.cpp: the lines fgets(1024) returns, and a position."""

    def __init__(self, path):
        raw = open(path, "rb").read().decode("latin-1")
        self.lines = []
        for line in raw.splitlines(keepends=True):
            while len(line) > 1023:
                self.lines.append(line[:1023])
                line = line[1023:]
            self.lines.append(line)
        self.pos = 0

    def _scan(self, title):
        while self.pos < len(self.lines):
            l = self.lines[self.pos]
            self.pos += 1
            c = l.find("%")
            if c >= 0:
                l = l[:c]
            t = l.find(title)
            if t >= 0:
                return l[t + len(title):]
        return None

    def readparam(self, title, wrap=True):
        r = self._scan(title)
        if r is None and wrap:
            self.pos = 0
            r = self._scan(title)
        return r

    def next_raw(self):
        if self.pos >= len(self.lines):
            return None
        self.pos += 1
        return self.lines[self.pos - 1]


def strnext(s: str, i: int, delim: str):
    """Index of the next word after a run of delimiters, None at end of string / newline."""
    found = False
    while i < len(s):
        if s[i] in delim:
            found = True
        if found and s[i] not in delim:
            break
        i += 1
    if i >= len(s) or s[i] == "\n":
        return None
    return i


def z_number(line: str) -> int:
    e0 = line[0] if len(line) > 0 else "\0"
    e1 = line[1] if len(line) > 1 else "\0"
    if e1.isdigit() and e1 != "0" or e1 in "\n\0\r":
        e1 = " "
    hit = EL_TABLE.find(e0 + e1) if e0 != "\0" else 0
    return hit // 2 + 1 if hit >= 0 else 0


def rot_matrix(px, py, pz):
    c, s = math.cos, math.sin
    return np.array([
        [c(pz) * c(py), c(pz) * s(py) * s(px) - s(pz) * c(px), c(pz) * s(py) * c(px) + s(pz) * s(px)],
        [s(pz) * c(py), s(pz) * s(py) * s(px) + c(pz) * c(px), s(pz) * s(py) * c(px) - c(pz) * s(px)],
        [-s(py), c(py) * s(px), c(py) * c(px)]], f64)


def _rot(M, u):
    """rotateVect's row-by-row products in the reference's summation order."""
    return np.array([M[r, 0] * u[0] + M[r, 1] * u[1] + M[r, 2] * u[2] for r in range(3)], f64)


def wavelength_A(kev: float) -> float:
    emass, hc = 510.99906, 12.3984244
    return hc / math.sqrt(kev * (2 * emass + kev))


def read_unit_cell(cfg, ncx, ncy, ncz, ctilt, xoff, yoff):
    """-> (Z int32[n], xyz float32[n,3] in Angstrom, dw float32[n], occ float32[n], box float32[3])"""
    f = ParFile(cfg)
    ncoord, scale = 0, 0.0
    Mm = np.zeros((3, 3), f64)
    r = f.readparam("Number of particles =")
    if r is not None:
        ncoord = _scan_d(r, 0)
    r = f.readparam("A =")
    if r is not None:
        scale = _atof(r)
    for a in range(3):
        for b in range(3):
            r = f.readparam(f"H0({a + 1},{b + 1}) =")
            if r is not None:
                Mm[a, b] = _atof(r)
    Mm *= scale
    if ncoord < 1:
        raise ValueError("ncoord = 0")
    f.pos = 0
    no_vel = f.readparam(".NO_VELOCITY.") is not None
    entry = 3
    r = f.readparam("entry_count =")
    if r is not None:
        entry = _scan_d(r, entry)
    if not no_vel:
        entry += 3
    off = 0 if no_vel else 3
    cell = [None] * ncoord
    mass, element = 28.0, 1
    for i in range(ncoord - 1, -1, -1):
        line = f.next_raw()
        if line is None:
            raise ValueError("number of atoms does not agree with atoms in file")
        nx = strnext(line, 0, " \t")
        if _atof(line) >= 1.0 and (nx is None or line[nx] == "#"):
            mass = _atof(line)
            element = z_number(f.next_raw())
            line = f.next_raw()
        s = 0
        while s < len(line) and line[s] in " \t":
            s += 1
        data = []
        for _ in range(entry):
            if s is None:
                raise ValueError("incomplete data line: " + line)
            data.append(_atof(line[s:]))
            s = strnext(line, s, " \t")
        dw = f32(0.45 * 28.0 / mass)
        occ = f32(1.0)
        if entry > 3 + off:
            dw = f32(data[3 + off])
        if entry > 4 + off:
            occ = f32(data[4 + off])
        cell[i] = (f32(data[2]), f32(data[1]), f32(data[0]), dw, occ, element)   # z, y, x like the struct
    cell.sort(key=lambda a: (a[0], a[1], a[2]))
    for i in range(ncoord - 1, -1, -1):
        if cell[i][4] < 1:
            raise ValueError("partial occupancy (reference: ran1 lottery)")
        if i > 0 and all(abs(float(cell[i][k]) - float(cell[i - 1][k])) < 1e-6 for k in range(3)):
            raise ValueError("shared site (reference: ran1 lottery)")
    n = ncoord * ncx * ncy * ncz
    Z = np.zeros(n, np.int32)
    frac = np.zeros((n, 3), f32)
    dwv, occv = np.zeros(n, f32), np.zeros(n, f32)
    for icx in range(ncx):
        for icy in range(ncy):
            for icz in range(ncz):
                j0 = (icz + icy * ncz + icx * ncy * ncz) * ncoord
                for i, (z, y, x, dw, occ, el) in enumerate(cell):
                    Z[j0 + i], dwv[j0 + i], occv[j0 + i] = el, dw, occ
                    frac[j0 + i] = (f32(x + f32(icx)), f32(y + f32(icy)), f32(z + f32(icz)))
    fx, fy, fz = (frac[:, k].astype(f64) for k in range(3))
    xyz = np.stack([Mm[0, c] * fx + Mm[1, c] * fy + Mm[2, c] * fz for c in range(3)], 1).astype(f32)
    bc = np.array([ncx / 2.0, ncy / 2.0, ncz / 2.0])
    ctr = np.array([Mm[0, c] * bc[0] + Mm[1, c] * bc[1] + Mm[2, c] * bc[2] for c in range(3)])
    M = rot_matrix(float(ctilt[0]), float(ctilt[1]), float(ctilt[2]))
    corners = []
    for icx in (0, ncx):
        for icy in (0, ncy):
            for icz in (0, ncz):
                u = np.array([Mm[0, c] * (icx - bc[0]) + Mm[1, c] * (icy - bc[1]) + Mm[2, c] * (icz - bc[2])
                              for c in range(3)])
                corners.append(_rot(M, u) + ctr)
    corners = np.array(corners)
    lo, hi = corners.min(0), corners.max(0)
    if any(float(t) != 0 for t in ctilt):
        u = xyz.astype(f64) - ctr
        xyz = (np.stack([M[r, 0] * u[:, 0] + M[r, 1] * u[:, 1] + M[r, 2] * u[:, 2] for r in range(3)], 1) + ctr).astype(f32)
    xyz = (xyz.astype(f64) - lo).astype(f32)
    box = (hi - lo).astype(f32)
    if xoff != 0 or yoff != 0:
        xyz[:, 0] = xyz[:, 0] + f32(xoff)
        xyz[:, 1] = xyz[:, 1] + f32(yoff)
    return Z, xyz, dwv, occv, box


def read_qsc(path, atoms_from_external=False, return_shift=False):
    """-> (Params after consitentParams, Z, xyz [m], DWF [m^2], occ), like fdes_oracle.read_cnf."""
    q = ParFile(path)
    pi = 3.1415926535897
    r = q.readparam("mode:")
    if r is None or "TEM" not in r:
        raise ValueError("FDES supports only TEM mode")
    q.readparam("print level:")
    q.readparam("save level:")
    r = q.readparam("filename:")
    if r is None:
        raise ValueError("no filename:")
    base = _word(r)
    if base.startswith('"'):
        base = r[r.find('"') + 1:]
        base = base[:base.find('"')]
    q.readparam("wavename:")
    ncx = _scan_d(q.readparam("NCELLX:") or "", 0)
    ncy = _scan_d(q.readparam("NCELLY:") or "", 0)
    ncz, celldiv = 0, 1
    r = q.readparam("NCELLZ:")
    if r is not None:
        a = _word(r)
        if "/" in a:
            a, d = a.split("/", 1)
            celldiv = _scan_d(d, 0)
        ncz = _scan_d(a, 0)

    def angle(key):
        v = f32(0)
        r = q.readparam(key)
        if r is not None:
            m = _NUM.match(r)
            if m:
                v = f32(float(m.group(1)))
                unit = _word(r[m.end():])
                if unit[:1].lower() == "d":
                    v = f32(f64(v) * (pi / 180.0))
        return v

    btx, bty = angle("Beam tilt X:"), angle("Beam tilt Y:")
    q.readparam("Tilt back:")
    ctilt = (angle("Crystal tilt X:"), angle("Crystal tilt Y:"), angle("Crystal tilt Z:"))
    r = q.readparam("Cube:")
    if r is not None:
        c = [_atof(t) for t in r.split()[:3]]
        if len(c) == 3 and all(v > 0 for v in c):
            raise ValueError("Cube: mode")
    q.readparam("Adjust cube size with tilt:")
    r = q.readparam("tds:")
    if r is not None and _word(r)[:1].lower() == "y":
        raise ValueError("tds: yes")
    q.readparam("temperature:")
    q.readparam("phonon-File:")
    pos_file = base if "." in base else base + ".cfg"
    if not pos_file.endswith(".cfg"):
        raise ValueError("only .cfg specimen files")
    cfg = pos_file if os.path.exists(pos_file) else os.path.join(os.path.dirname(str(path)), pos_file)
    xoff = _scan_g(q.readparam("xOffset:") or "", f32(0))
    yoff = _scan_g(q.readparam("yOffset:") or "", f32(0))
    Z, xyzA, dw, occ, box = read_unit_cell(cfg, ncx, ncy, ncz, ctilt, xoff, yoff)

    r = q.readparam("nx:")
    if r is None:
        raise ValueError("no nx:")
    nx = _scan_d(r, 0)
    r = q.readparam("ny:")
    ny = _scan_d(r, 0) if r is not None else nx
    resX = _scan_g(q.readparam("resolutionX:") or "", f32(0))
    resY = _scan_g(q.readparam("resolutionY:") or "", f32(0))
    r = q.readparam("v0:")
    if r is None:
        raise ValueError("no v0:")
    v0 = _scan_g(r, f32(0))
    center = 0
    r = q.readparam("center slices:")
    if r is not None:
        center = int(_word(r)[:1].lower() == "y")
    thick, slices = f32(0), 0
    r = q.readparam("slice-thickness:")
    if r is not None:
        thick = _scan_g(r, f32(0))
        r = q.readparam("slices:")
        if r is not None:
            slices = _scan_d(r, 0)
        else:
            slices = int(f64(f32(box[2] / f32(f32(celldiv) * thick))) + 0.99)
        slices += center
    else:
        r = q.readparam("slices:")
        if r is not None:
            slices = _scan_d(r, 0)
            thick = f32(box[2] / f32(celldiv)) if (slices == 1 and celldiv == 1) else f32(box[2] / f32(celldiv * slices))
    if slices == 0:
        raise ValueError("Number of slices = 0")
    q.readparam("slices between outputs:")
    q.readparam("zOffset:")
    if resX == 0:
        resX = f32(f64(box[0]) / nx)
    if resY == 0:
        resY = f32(f64(box[1]) / ny)
    for key in ("periodicXY:", "periodicZ:", "bandlimit f_trans:", "read potential:", "save potential:",
                "save projected potential:", "plot V(r)*r:", "one time integration:", "potential3D:",
                "Runs for averaging:", "Store TDS diffr. patt. series:", "potential progress interval:",
                "dE/E:", "dI/I:", "dV/V:", "Cc:"):
        q.readparam(key)
    r = q.readparam("Cs:")
    if r is None:
        raise ValueError("no Cs:")
    Cs = f32(f64(_scan_g(r, f32(0))) * 1.0e7)
    C5 = f32(0)
    r = q.readparam("C5:")
    if r is not None:
        C5 = f32(f64(_scan_g(r, f32(0))) * 1.0e7)
    scherzer = lambda k: f32(-f64(f32(math.sqrt(k * f64(Cs) * wavelength_A(f64(v0))))))
    df0 = scherzer(1.5)
    r = q.readparam("defocus:")
    if r is not None:
        a = _word(r)[:1].lower()
        if a == "s":
            df0 = scherzer(1.5)
        elif a == "o":
            df0 = scherzer(1.0)
        else:
            df0 = f32(10.0 * f64(_scan_g(r, df0)))
    astig = f32(10.0 * f64(_scan_g(q.readparam("astigmatism:") or "", f32(0))))
    astig_angle = f32(f64(_scan_g(q.readparam("astigmatism angle:") or "", f32(0))) * (pi / 180.0))
    r = q.readparam("alpha:")
    if r is None:
        raise ValueError("no alpha:")
    alpha = _scan_g(r, f32(0))

    p = Params()
    p.pi = f32(3.1415927)
    p.n3 = 1
    p.tiltspec, p.tiltbeam, p.defoci = np.zeros(2, f32), np.zeros(2, f32), np.zeros(1, f32)
    p.n1, p.n2 = nx, ny
    p.dn1, p.dn2 = nx // 2, ny // 2
    p.m3 = slices
    p.d1 = f32(f64(resX) * 1e-10)
    p.d2 = f32(f64(resY) * 1e-10)
    p.d3 = f32(f64(thick) * 1e-10)
    p.subSlTh = f32(f64(thick) * 1e-10 / 10)
    p.tilt_off = ctilt
    p.tiltbeam[0], p.tiltbeam[1] = btx, bty
    p.E0 = f32(f64(v0) * 1e3)
    p.illangle = f32(f64(alpha) / 1e3)
    p.ab0["A1"] = f32(f64(astig) * 1e-9)
    p.ab1["A1"] = f32(f64(astig_angle) * 1e-9)
    p.ab0["C1"] = f32(f64(df0) * 1e-10)
    p.ab0["C3"] = f32(f64(Cs) * 1e-10)
    p.ab0["C5"] = f32(f64(C5) * 1e-3)
    for key, attr, kind in (("cal_mode:", "mode", "d"), ("focus_spread:", "defocspread", "g"),
                            ("objective_aperture:", "ObjAp", "g"), ("pixel_dose:", "pD", "g"),
                            ("absorptive_potential_factor:", "imPot", "g"), ("mtf_a:", "mtfa", "g"),
                            ("mtf_b:", "mtfb", "g"), ("mtf_c:", "mtfc", "g"), ("frozen_phonons:", "frPh", "d")):
        r = q.readparam(key)
        if r is not None:
            setattr(p, attr, (_scan_d if kind == "d" else _scan_g)(r, getattr(p, attr)))
    xyz = np.zeros((0, 3), f32)
    shift = np.zeros(3, f32)
    if not atoms_from_external:
        xyz = (xyzA.astype(f64) * 1e-10).astype(f32)
        dw = (dw.astype(f64) * 1e-20).astype(f32)
        mx = np.maximum(xyz.max(0), f32(0))
        mn = np.minimum(xyz.min(0), f32(1))
        shift = ((mx - mn) / f32(2)).astype(f32)
        xyz = (xyz - shift).astype(f32)
    else:
        Z, dw, occ = Z[:0], dw[:0], occ[:0]
    consistent_params(p)
    if return_shift:
        return p, Z, xyz, dw, occ, shift
    return p, Z, xyz, dw, occ


def read_qsc_scan(path):
    """STEM raster and detectors of a .qsc: the keys of src/rwQsc.cu:444-466 and :698-735.
    -> (positions float32 [nx, ny, 2] in metres in the frame of the centred atoms, detectors [ndet, 2] mrad)."""
    shift = read_qsc(path, return_shift=True)[5]
    q = ParFile(path)

    def need(key, scan):
        r = q.readparam(key)
        v = scan(r, None) if r is not None else None
        if v is None:
            raise ValueError("STEM scan needs " + key)
        return v

    xs, xe, nx = need("scan_x_start:", _scan_g), need("scan_x_stop:", _scan_g), max(1, need("scan_x_pixels:", _scan_d))
    ys, ye, ny = need("scan_y_start:", _scan_g), need("scan_y_stop:", _scan_g), max(1, need("scan_y_pixels:", _scan_d))
    dx, dy = f32(f32(xe - xs) / f32(nx)), f32(f32(ye - ys) / f32(ny))
    xy = np.zeros((nx, ny, 2), f32)
    for ix in range(nx):
        for iy in range(ny):
            x, y = f32(xs + f32(f32(ix) * dx)), f32(ys + f32(f32(iy) * dy))
            xy[ix, iy, 0] = f32(f32(f64(x) * 1e-10) - shift[0])
            xy[ix, iy, 1] = f32(f32(f64(y) * 1e-10) - shift[1])
    det = []
    q.pos = 0
    while True:
        r = q.readparam("detector:", wrap=False)
        if r is None:
            break
        m1 = _NUM.match(r)
        m2 = _NUM.match(r[m1.end():]) if m1 else None
        if m1 and m2:
            det.append((f32(float(m1.group(1))), f32(float(m2.group(1)))))
    return xy, np.array(det, f32).reshape(-1, 2)
