"""CPU: host-side logic of the product (no compute calls): the C-ABI library loads and exports
every symbol include/fdes_b200.h declares, the C++ .cnf reader agrees bit-for-bit with the oracle's
restatement of the reference reader, and the product never routes through the oracle."""
import ctypes
import pathlib
import re
import subprocess

import numpy as np
import pytest

from conftest import CASES, DATA, ROOT


def test_library_exports_declared_symbols(fb):
    lib = fb.load_library()
    names = fb.declared_symbols()
    assert "FDES" in names and len(names) >= 20
    for n in names:
        assert getattr(lib, n) is not None
    # the name the reference's Python binding loads (Python/pyFDES.py:36) resolves to the same library
    alias = ctypes.CDLL(str(fb.LIB_PATH.parent / "libFDES_SHARED_LIB.so"))
    assert alias.FDES is not None


def test_version_and_error_channel(fb):
    lib = fb.load_library()
    assert lib.fdes_b200_version() >= 100
    assert lib.fdes_b200_parse_cnf(b"/nonexistent/x.cnf", None, None, None, None, 0) == -1
    assert b"cannot read" in lib.fdes_b200_last_error()


@pytest.mark.parametrize("case", CASES)
def test_cnf_reader_matches_oracle_reader(case, fb, orc):
    r = fb.parse_cnf(DATA / f"{case}.cnf")
    p, Z, xyz, dwf, occ = orc.read_cnf(str(DATA / f"{case}.cnf"))
    orc.set_sub_slices(p, orc.sub_slice_ratio(p.d3, p.subSlTh))
    for k, v in dict(n1=p.n1, n2=p.n2, n3=p.n3, m1=p.m1, m2=p.m2, m3=p.m3, nAt=len(Z), frPh=p.frPh, mode=p.mode,
                     nZ=len(orc.list_of_elements(Z))).items():
        assert r[k] == v, k
    for k, v in dict(lam=p.lam, sigma=p.sigma, gamma=p.gamma, d1=p.d1, d2=p.d2, d3=p.d3, E0=p.E0, imPot=p.imPot).items():
        assert np.float32(r[k]) == np.float32(v), k       # bit-exact float32
    np.testing.assert_array_equal(r["atoms"][:, 0].astype(np.int32), Z)
    np.testing.assert_array_equal(r["atoms"][:, 1:4], xyz)
    np.testing.assert_array_equal(r["atoms"][:, 4], dwf)
    np.testing.assert_array_equal(r["atoms"][:, 5], occ)
    np.testing.assert_array_equal(r["tiltspec"].ravel(), p.tiltspec[: 2 * p.n3])
    np.testing.assert_array_equal(r["tiltbeam"].ravel(), p.tiltbeam[: 2 * p.n3])
    np.testing.assert_array_equal(r["defoci"], p.defoci[: p.n3])


def test_trailing_newline_duplicates_last_atom(fb, tmp_path):
    """Reference reader quirk (src/paramStructure.cu:1019-1077): a file ending in "atom: ...\\n"
    yields its last atom twice."""
    src = (DATA / "tem64.cnf").read_text()
    a = tmp_path / "a.cnf"
    b = tmp_path / "b.cnf"
    a.write_text(src.rstrip("\n"))
    b.write_text(src.rstrip("\n") + "\n")
    ra, rb = fb.parse_cnf(a), fb.parse_cnf(b)
    assert rb["nAt"] == ra["nAt"] + 1
    np.testing.assert_array_equal(rb["atoms"][-1], ra["atoms"][-1])


def test_sub_slicing(fb, tmp_path):
    from fdes_b200 import specimens
    specimens.write_cnf(tmp_path / "s.cnf", image_size=32, border_size=16, slices=12, pixel_size=0.25e-10,
                        slice_thickness=2.1e-10, sub_slice_thickness=0.2e-10, atoms=specimens.au_cuboctahedron(1))
    r = fb.parse_cnf(tmp_path / "s.cnf")
    assert r["m3"] == 12 * 11 and r["nAt"] == 13      # ceil(2.1/0.2) = 11 (src/crystalMaker.cu:720-733)
    assert abs(float(r["d3"]) - 2.1e-10 / 11) < 1e-16


def test_specimen_generators():
    from fdes_b200 import specimens
    si = specimens.si001_slab()
    assert si.shape == (11552, 6)                      # ExampleSpecimens/Si_001_11k_cnf
    assert np.allclose(si[:, 1:3].max(0), 5.0912906e-9, rtol=1e-6) and np.allclose(si[:, 3].max(), 1.0182581e-9, rtol=1e-6)
    assert specimens.au_cuboctahedron().shape == (309, 6)   # ExampleSpecimens/Au_cubeoctahedron_*
    assert specimens.srtio3_slab(2, 2, 3).shape == (60, 6)


def test_bench_workloads_are_the_baseline_shapes(fb, tmp_path):
    """The grids / slice counts of the bench workloads (bench.py WORKLOADS) as the library's reader sees them."""
    import importlib.util
    from fdes_b200 import specimens
    spec = importlib.util.spec_from_file_location("bench", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    want = {"srtio3_800": (800, 400, 3), "si001_1024": (1024, 11, 1), "au_2048": (2048, 12, 1), "slab_4096": (4096, 20, 3)}
    for name, (m, slices, nz) in want.items():
        assert name in bench.WORKLOADS and name in bench.DEFAULT_CONFIGS_PER_STEP and name in bench.REF_SAMPLE_CONFIGS
        cnf = tmp_path / f"{name}.cnf"
        getattr(specimens, bench.WORKLOADS[name][0])(cnf, frozen_phonons=2)
        r = fb.parse_cnf(cnf)
        assert (r["m1"], r["m2"], r["m3"], r["nZ"], r["frPh"]) == (m, m, slices, nz, 2)


def test_no_gpu_fails_loudly(fb):
    """There is no CPU fallback: without a CUDA device every computing entry point refuses."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(fb.FdesError, match="no CUDA device|CUDA"):
        fb.Simulation(DATA / "tem64.cnf")
    with pytest.raises(fb.FdesError):
        fb.fft2d(np.zeros((64, 64), np.complex64))


def test_product_never_touches_the_oracle():
    pat = re.compile(r"fdes_oracle|oracle/|_ref")
    for f in list((ROOT / "fdes_b200").glob("*.py")) + list((ROOT / "fdes_b200" / "csrc").glob("*")):
        if f.suffix in (".py", ".cu", ".cuh", ".cpp", ".h"):
            assert not pat.search(f.read_text()), f


def test_cli_usage():
    exe = ROOT / "fdes_b200" / "bin" / "FDES"
    assert exe.exists(), "build with __graft_entry__.build()"
    r = subprocess.run([str(exe), "--help"], capture_output=True, text=True)
    assert "--input_name" in r.stderr and "--print_level" in r.stderr


def test_shard_ranges():
    from fdes_b200.distributed import shard_range
    for count in (1, 3, 32, 64, 7):
        for world in (1, 2, 4, 8):
            r = [shard_range(count, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == count
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_used_cnf_atom_table_is_formatted_like_printf(fb, tmp_path):
    """writeConfig prints the atom table with "%i %14.8g %14.8g %14.8g %14.8g %14.8g \\n"
    (src/paramStructure.cu:480-484); the product formats it with std::to_chars on a few threads --
    every line must be what printf would have written, over 35 decades and the awkward values."""
    from fdes_b200 import specimens
    rng = np.random.default_rng(0)
    vals = np.concatenate([
        np.array([0.0, -0.0, 1.0, -1.5, 1e-10, 5.3837e-21, 123456789.0, 1.23456785e-5, 9.9999999e9, 3.4e38, 1e-38, 0.1,
                  2 / 3, 99999999.5, 0.00001, 1e-5, 123456.78, 0.000123456785, 1e-4, 9.9999995e-5], np.float32),
        (rng.standard_normal(12000) * 10.0 ** rng.integers(-25, 10, 12000)).astype(np.float32)])
    v = vals[: len(vals) // 5 * 5].reshape(-1, 5)
    atoms = np.column_stack([np.full(len(v), 14, np.float32), v]).astype(np.float32)
    specimens.write_cnf(tmp_path / "adv.cnf", image_size=32, border_size=16, slices=2, pixel_size=0.25e-10,
                        slice_thickness=2e-10, atoms=atoms)
    lib = fb.load_library()
    assert lib.fdes_b200_write_used_cnf(str(tmp_path / "adv.cnf").encode(), str(tmp_path / "used.txt").encode()) == 0
    a = fb.parse_cnf(tmp_path / "adv.cnf")["atoms"]
    lines = (tmp_path / "used.txt").read_text().splitlines()
    i0 = next(i for i, l in enumerate(lines) if l.startswith("# Atomic no.")) + 1
    assert len(a) >= 2400 and len(lines) - i0 == len(a)
    for k, line in enumerate(lines[i0:]):
        assert line == "%i %14.8g %14.8g %14.8g %14.8g %14.8g " % (int(a[k, 0]), *[float(x) for x in a[k, 1:6]]), k


def test_gpu_list_syntax(fb):
    """FDES_B200_GPUS / --gpus values (no device needed: the count form is not clamped without one)."""
    assert fb.parse_gpu_list(None, 3) == [3]
    assert fb.parse_gpu_list("", 1) == [1]
    assert fb.parse_gpu_list("0,2,5") == [0, 2, 5]
    assert fb.parse_gpu_list("1,") == [1]
    import torch
    if not torch.cuda.is_available():
        assert fb.parse_gpu_list("4", 2) == [2, 3, 4, 5]
