"""CPU, world_size 2 over gloo: the multi-GPU driver (fdes_b200/distributed.py) -- shard the
frozen-phonon configurations, all-reduce the partial intensity / exit-wave sums, apply the detector
tail once.  The CUDA session is replaced by a stand-in with the same interface that computes its
shard with the oracle (tests may use the oracle as a checker; the product never does), so this
covers exactly the host logic that runs at N > 1: ranges, RNG burn-in, reduction, scaling."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import DATA, ROOT, TOL_INTENSITY, TOL_WAVE, rel_l2


class OracleSession:
    """Session interface of fdes_b200.Simulation on top of the numpy oracle, for one rank."""

    def __init__(self, cnf, rank, world):
        import fdes_oracle as orc
        from fdes_b200.distributed import shard_range
        self.orc = orc
        p, self.Z, self.xyz, self.dwf, self.occ = orc.read_cnf(str(cnf))
        orc.set_sub_slices(p, orc.sub_slice_ratio(p.d3, p.subSlTh))
        self.p = p
        self.n1, self.n2, self.n3, self.m1, self.m2 = p.n1, p.n2, p.n3, p.m1, p.m2
        self.count = max(1, p.frPh)
        self.j0, self.j1 = shard_range(self.count, rank, world)
        self.rng = orc.Xorwow(1, 3 * len(self.Z)) if p.frPh > 0 else None
        self.burn = self.j0          # normals of the configurations owned by lower ranks
        self.I = self.E = None

    def set_accumulators(self, i_ptr, e_ptr=0):
        n = self.m1 * self.m2
        self.I = np.ctypeslib.as_array((ctypes.c_float * n).from_address(i_ptr))
        self.E = np.ctypeslib.as_array((ctypes.c_float * (2 * n)).from_address(e_ptr)).view(np.complex64) if e_ptr else None

    def run_k(self, k):
        orc, p = self.orc, self.p
        f32 = np.float32
        self.I[:] = 0
        if self.E is not None:
            self.E[:] = 0
        xyzTO = orc.tilt_coordinates(self.xyz, *p.tilt_off)
        xyz_k = orc.tilt_coordinates(xyzTO, p.tiltspec[2 * k], p.tiltspec[2 * k + 1], 0.0)
        mask = orc.band_mask(p)
        P = orc.fresnel_propagator(p, mask)
        Zl = orc.list_of_elements(self.Z)
        alpha = f32(f32(1) / f32(self.count))
        for _ in range(self.burn):
            orc.atom_jitter(xyz_k, self.dwf, self.rng)
        self.burn = 0
        for j in range(self.j0, self.j1):
            psi = orc.incoming_wave(p, k, mask)
            xyzFP = orc.atom_jitter(xyz_k, self.dwf, self.rng) if p.frPh > 0 else xyz_k
            bins = orc.bin_atoms(xyzFP, p)
            for s in range(p.m3):
                V = orc.phase_grating(s, self.Z, Zl, xyzFP, self.occ, p.imPot, p, bins)
                psi = orc.forward_propagation(psi, V, p, P, mask)
            if self.E is not None:
                self.E += (alpha * psi).astype(np.complex64).ravel()
            q = orc.apply_lens_function(psi, p, k) if p.mode == 0 else None
            inten = (q.real ** 2 + q.imag ** 2).astype(f32) if p.mode == 0 else orc.diffraction_pattern(psi, p, k, mask)
            self.I += (alpha * inten).astype(f32).ravel()

    def finish_k(self, k):
        img = self.orc.add_noise_and_mtf(self.I.reshape(self.m2, self.m1).astype(np.complex64), self.p, k)
        ew = self.E.reshape(self.m2, self.m1).copy() if self.E is not None else None
        return img, ew

    def close(self):
        pass


def _worker(rank, world, port, cnf, out):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fdes_b200.distributed import simulate_sharded
    img, ew = simulate_sharded(lambda r, w: OracleSession(cnf, r, w), want_exitwave=True, device=torch.device("cpu"))
    np.savez(f"{out}/rank{rank}.npz", img=img, ew=ew)
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("case", ["phonon64", "tilt64"])
def test_two_ranks_equal_single_process(case, tmp_path, oracle_runs):
    mp.spawn(_worker, args=(2, _free_port(), str(DATA / f"{case}.cnf"), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    np.testing.assert_array_equal(r0["img"], r1["img"])       # every rank ends with the same result
    np.testing.assert_array_equal(r0["ew"], r1["ew"])
    ref, _ = oracle_runs(case)                                 # single-process oracle, all configurations
    assert rel_l2(r0["ew"], ref.exitwave) < TOL_WAVE
    assert rel_l2(r0["img"], ref.image) < TOL_INTENSITY


class FakeStemSession:
    """stem_scan stand-in: a deterministic function of the probe position (checks ranges / ordering)."""
    def stem_scan(self, pos, det, k=0):
        pos = np.asarray(pos, np.float32)
        return np.stack([pos[:, 0] * 3 + pos[:, 1], pos[:, 0] - 2 * pos[:, 1]], 1).astype(np.float32), 1.0

    def close(self):
        pass


def _stem_worker(rank, world, port, out):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fdes_b200.distributed import stem_scan_sharded
    pos = np.arange(2 * 37, dtype=np.float32).reshape(37, 2)      # 37 probes: uneven split
    sig = stem_scan_sharded(lambda r, w: FakeStemSession(), pos, np.array([[0, 10], [10, 20]], np.float32),
                            device=torch.device("cpu"))
    np.save(f"{out}/stem{rank}.npy", sig)
    dist.destroy_process_group()


def test_stem_positions_are_sharded_and_gathered(tmp_path):
    mp.spawn(_stem_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "stem0.npy"), np.load(tmp_path / "stem1.npy")
    pos = np.arange(2 * 37, dtype=np.float32).reshape(37, 2)
    want, _ = FakeStemSession().stem_scan(pos, None)
    np.testing.assert_array_equal(a, want)
    np.testing.assert_array_equal(b, want)
