"""CPU: the numpy oracle against the golden vectors recorded from the unmodified reference
(tests/golden/*.npz, tools/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from conftest import CASES, DATA, TOL_INTENSITY, TOL_WAVE, golden, rel_l2


@pytest.mark.parametrize("case", CASES)
def test_reader_and_derived_parameters(case, orc):
    g, meta = golden(case)
    p, Z, xyz, dwf, occ = orc.read_cnf(str(DATA / f"{case}.cnf"))
    orc.set_sub_slices(p, orc.sub_slice_ratio(p.d3, p.subSlTh))
    assert (p.n1, p.n2, p.n3, p.m1, p.m2, p.m3) == tuple(int(meta[k]) for k in ("n1", "n2", "n3", "m1", "m2", "m3"))
    assert len(Z) == int(meta["nAt"])            # includes the trailing-newline duplicate (dup64)
    assert p.mode == int(meta["mode"]) and p.frPh == int(meta["frPh"])
    # meta.txt was printed with %.9g: float32 values round-trip exactly
    for key, val in (("lambda", p.lam), ("sigma", p.sigma), ("gamma", p.gamma), ("d1", p.d1), ("d3", p.d3)):
        assert np.float32(meta[key]) == np.float32(val), key
    # first configuration without phonons is the file's coordinates after the tilts
    if p.frPh == 0 and not np.any(p.tiltspec) and not any(p.tilt_off):
        np.testing.assert_array_equal(g["xyz_cfg"][0], xyz)


def test_xorwow_jitter_matches_curand(orc):
    """atomJitter_d with curand_init(1, i, 0) / curand_normal (src/crystalMaker.cu:28-48)."""
    g, _ = golden("phonon64")
    p, Z, xyz, dwf, occ = orc.read_cnf(str(DATA / "phonon64.cnf"))
    rng = orc.Xorwow(1, 3 * len(Z))
    for j in range(p.frPh):
        got = orc.atom_jitter(xyz, dwf, rng)
        ref = g["xyz_cfg"][j]
        # displacements are ~1e-11 m on coordinates of ~1e-10..1e-9 m: compare the displacement itself
        assert rel_l2(got - xyz, ref - xyz) < 1e-5, j


@pytest.mark.parametrize("case", CASES)
def test_potential_and_first_slices(case, orc):
    g, meta = golden(case)
    p, Z, xyz, dwf, occ = orc.read_cnf(str(DATA / f"{case}.cnf"))
    orc.set_sub_slices(p, orc.sub_slice_ratio(p.d3, p.subSlTh))
    xyz0 = g["xyz_cfg"][0]
    Zl = orc.list_of_elements(Z)
    mask = orc.band_mask(p)
    psi = orc.incoming_wave(p, 0, mask)
    assert rel_l2(psi, g["psi_in"]) < TOL_WAVE
    for s in range(g["V"].shape[0]):
        V = orc.phase_grating(s, Z, Zl, xyz0, occ, p.imPot, p)
        if np.linalg.norm(g["V"][s]) > 0:
            assert rel_l2(V, g["V"][s]) < TOL_WAVE, s
        else:
            assert not V.any()
        psi = orc.forward_propagation(psi, V, p, None, mask)
        assert rel_l2(psi, g["psi_s"][s]) < TOL_WAVE, s


@pytest.mark.parametrize("case", CASES)
def test_full_run(case, orc, oracle_runs):
    g, meta = golden(case)
    res, (p, Z, xyz, dwf, occ) = oracle_runs(case)
    assert rel_l2(res.exitwave, g["exitwave"]) < TOL_WAVE
    assert rel_l2(res.image, g["image"]) < TOL_INTENSITY
    # the trace replay of k = 0 agrees with the stock driver of the reference itself
    assert rel_l2(g["exitwave_avg_k0"], g["exitwave"][0]) < TOL_WAVE
    assert rel_l2(g["J_k0"], g["image"][0]) < TOL_INTENSITY


def test_edge_case_bins(orc):
    """Acceptance window 1 < x1 < m-2 and slice range (squareAtoms_d, src/crystalMaker.cu:85-92):
    the edge64 atoms were placed on the borders on purpose."""
    p, Z, xyz, dwf, occ = orc.read_cnf(str(DATA / "edge64.cnf"))
    i1, i2, i3, r1, r2, ok = orc.bin_atoms(xyz, p)
    inside = ok & (i3 >= 0) & (i3 < p.m3)
    # atoms 3 (x1 == 1), 5 (x1 == m-2), 8 (z above), 9 (z below) are dropped
    assert inside.tolist() == [True, True, True, False, True, False, True, True, False, False, True, True]
    assert i1[0] == 10 and i2[0] == 20 and r1[0] == 0 and r2[0] == 0


def test_empty_specimen_is_identity(orc):
    """No atom inside the acceptance window: psi stays the plane wave, image == 1 (the MTF has
    unit DC gain)."""
    p, Z, xyz, dwf, occ = orc.read_cnf(str(DATA / "tem64.cnf"))
    far = xyz.copy()
    far[:, 2] = 1e-6   # all atoms above the slab -> dropped
    res = orc.build_measurements(p, Z, far, dwf, occ)
    assert np.allclose(res.exitwave, 1.0, atol=1e-6)
