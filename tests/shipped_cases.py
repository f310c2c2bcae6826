"""The reference's shipped example inputs, run AS SHIPPED (tests/data/shipped/ holds unmodified copies of
/root/reference/bin and /root/reference/ExampleSpecimens: they are input data, not code) plus one
variant each with `pixel_dose: 0` (deterministic detector tail).  Shared by tools/make_shipped_golden.py
(runs the unmodified reference on a GPU box -> tests/golden/shipped/*.npz) and tests/test_shipped_inputs.py.

Golden content per case (reductions keep the repository small): the image of the measurements K_KEEP,
mean and L2 norm of every image, a central crop of the coherent exit wave of the same measurements and
the sum of |psi|^2 of every exit wave."""
import pathlib
import re
import shutil

import numpy as np

SHIPPED = pathlib.Path(__file__).resolve().parent / "data" / "shipped"

# name -> (file relative to SHIPPED, pixel dose override or None = as shipped)
CASES = {
    "bin_test_qsc": ("bin/test.qsc", None),                       # CBED (cal_mode 2), pixel_dose 10: test.qsc:80,87
    "bin_test_qsc_dose0": ("bin/test.qsc", 0.0),
    "bin_dataFDES": ("bin/dataFDES.cnf", None),                   # Au-309, 320^2, 25 tilts, dose 100
    "bin_dataFDES_dose0": ("bin/dataFDES.cnf", 0.0),
    "au_particle_cnf": ("ExampleSpecimens/Au_cubeoctahedron_cnf/dataFDES_Auparticle.cnf", None),
    "au_particle_cnf_dose0": ("ExampleSpecimens/Au_cubeoctahedron_cnf/dataFDES_Auparticle.cnf", 0.0),
    "si_001_11k": ("ExampleSpecimens/Si_001_11k_cnf/dataFDES_11k.cnf", None),   # 1000^2, 205 slices, dose 0 as shipped
    "srtio3_qsc": ("ExampleSpecimens/SrTiO3_qsc/SrTiO3.qsc", None),
    "srtio3_qsc_dose0": ("ExampleSpecimens/SrTiO3_qsc/SrTiO3.qsc", 0.0),
}
CROP = 96          # half-width of the central exit-wave crop
# focused probe through 400 sub-slices: the goldens also carry a float64 evaluation of the model
DEEP_PROBE_CASES = ("bin_test_qsc", "bin_test_qsc_dose0", "srtio3_qsc", "srtio3_qsc_dose0")


def stage(case: str, workdir: pathlib.Path) -> pathlib.Path:
    """Copy the input (and the .cfg a .qsc names) into workdir, apply the dose override, return the path."""
    rel, dose = CASES[case]
    src = SHIPPED / rel
    workdir.mkdir(parents=True, exist_ok=True)
    for f in src.parent.iterdir():
        if f.suffix == ".cfg":
            shutil.copy(f, workdir / f.name)
    dst = workdir / src.name
    raw = src.read_bytes()
    if dose is not None:
        raw, n = re.subn(rb"(?m)^(pixel_dose:\s*)[-+0-9.eE]+", lambda m: m.group(1) + repr(float(dose)).encode(), raw)
        assert n == 1, f"{src}: pixel_dose line not found"
    dst.write_bytes(raw)
    return dst


def k_keep(n3: int):
    return sorted({0, n3 // 2, n3 - 1})


def reduce(image: np.ndarray, exitwave: np.ndarray) -> dict:
    """image [n3][n2][n1] float32, exitwave [n3][m2][m1] complex64 -> the golden reductions."""
    n3 = image.shape[0]
    ks = k_keep(n3)
    m2, m1 = exitwave.shape[1:]
    c2, c1 = m2 // 2, m1 // 2
    h = min(CROP, c1, c2)
    return {
        "k_keep": np.array(ks, np.int32),
        "image_keep": np.ascontiguousarray(image[ks]),
        "image_mean": image.reshape(n3, -1).mean(axis=1, dtype=np.float64),
        "image_norm": np.sqrt((image.reshape(n3, -1).astype(np.float64) ** 2).sum(axis=1)),
        "ew_crop": np.ascontiguousarray(exitwave[ks][:, c2 - h:c2 + h, c1 - h:c1 + h]),
        "ew_power": (np.abs(exitwave.reshape(n3, -1).astype(np.complex128)) ** 2).sum(axis=1),
        "shape": np.array([n3, image.shape[1], image.shape[2], m2, m1], np.int32),
    }
