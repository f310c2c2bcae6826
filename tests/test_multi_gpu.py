"""GPU, world size 2 over NCCL (needs two B200s: `gpurun --gpus 2 -- python -m pytest tests -m gpu`):
frozen-phonon configurations sharded across ranks through fdes_b200.distributed.simulate_sharded,
compared with the single-GPU result of the same library and with the golden vectors of the reference.
Skipped when fewer than two GPUs are visible."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import DATA, ROOT, TOL_INTENSITY, TOL_WAVE, golden, rel_l2

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, cnf, out):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    import fdes_b200
    from fdes_b200.distributed import simulate_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    img, ew = simulate_sharded(lambda r, w: fdes_b200.Simulation(cnf, gpu_index=rank, rank=r, world=w, want_exitwave=True),
                               want_exitwave=True)
    np.savez(f"{out}/rank{rank}.npz", img=img, ew=ew)
    dist.destroy_process_group()


def test_two_gpus_match_one(tmp_path, fb):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cnf = str(DATA / "phonon64.cnf")
    mp.spawn(_worker, args=(2, port, cnf, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    np.testing.assert_array_equal(r0["img"], r1["img"])
    np.testing.assert_array_equal(r0["ew"], r1["ew"])
    with fb.Simulation(cnf, want_exitwave=True) as sim:
        img, ew = sim.simulate()
    assert rel_l2(r0["ew"], ew) < 1e-6 and rel_l2(r0["img"], img) < 1e-6
    g, _ = golden("phonon64")
    assert rel_l2(r0["ew"], g["exitwave"]) < TOL_WAVE
    assert rel_l2(r0["img"], g["image"]) < TOL_INTENSITY


# ---- one process, several engines (fdes_b200_open_multi) ---------------------------------------
# The engines of a multi-GPU session may also share a device: gpus=[0, 0] exercises the whole
# sharding / reduction path of the library on a one-GPU box; with two GPUs the same tests run over
# NVLink peer mappings.
def _gpu_pairs():
    import torch
    return [[0, 0]] + ([[0, 1]] if torch.cuda.device_count() >= 2 else [])


def test_configs_sharded_over_engines_match_one(fb):
    cnf = str(DATA / "phonon64.cnf")
    with fb.Simulation(cnf, want_exitwave=True) as sim:
        img, ew = sim.simulate()
    g, _ = golden("phonon64")
    for gpus in _gpu_pairs():
        with fb.Simulation(cnf, want_exitwave=True, gpus=gpus) as sim:
            assert sim.num_gpus == 2
            img2, ew2 = sim.simulate()
            img3, ew3 = sim.simulate()     # a second run of the same session: streams are re-seeded
        np.testing.assert_array_equal(img2, img3)
        np.testing.assert_array_equal(ew2, ew3)
        assert rel_l2(ew2, ew) < 1e-6 and rel_l2(img2, img) < 1e-6, gpus
        assert rel_l2(ew2, g["exitwave"]) < TOL_WAVE
        assert rel_l2(img2, g["image"]) < TOL_INTENSITY


def test_tilt_series_dealt_out_is_bit_identical(fb):
    cnf = str(DATA / "tilt64.cnf")
    with fb.Simulation(cnf, want_exitwave=True) as sim:
        assert sim.n3 > 1
        img, ew = sim.simulate()
    for gpus in _gpu_pairs():
        with fb.Simulation(cnf, want_exitwave=True, gpus=gpus) as sim:
            assert sim.num_gpus == 2
            img2, ew2 = sim.simulate()
        np.testing.assert_array_equal(img2, img)     # no reduction: every k is computed by one engine
        np.testing.assert_array_equal(ew2, ew)


def test_more_gpus_than_units_uses_fewer_engines(fb):
    with fb.Simulation(str(DATA / "tem64.cnf"), gpus=[0, 0, 0]) as sim:    # 1 configuration, 1 measurement
        assert sim.num_gpus == 1
        img, _ = sim.simulate()
    with fb.Simulation(str(DATA / "tem64.cnf")) as sim:
        ref, _ = sim.simulate()
    np.testing.assert_array_equal(img, ref)


def test_stem_probes_dealt_out(fb):
    cnf = str(DATA / "cbed64.cnf")
    pos = (np.stack(np.meshgrid(np.arange(3), np.arange(4), indexing="ij"), -1).reshape(-1, 2) - 1.0).astype(np.float32) * 2e-11
    det = np.array([[0.0, 30.0], [30.0, 300.0]], np.float32)
    with fb.Simulation(cnf, batch=4) as sim:
        ref, _ = sim.stem_scan(pos, det)
    for gpus in _gpu_pairs():
        with fb.Simulation(cnf, batch=4, gpus=gpus) as sim:
            assert sim.num_gpus == 2
            sig, ms = sim.stem_scan(pos, det)
        assert ms > 0
        np.testing.assert_allclose(sig, ref, rtol=2e-6, atol=0)


def test_FDES_on_two_devices_in_one_process_and_env_gpus(tmp_path, fb, monkeypatch):
    """Function attributes (shared-memory opt-in) are per device: FDES(gpu_Index=1) after
    FDES(gpu_Index=0) in one process must work; FDES_B200_GPUS shards the same call."""
    import torch
    cnf = DATA / "phonon64.cnf"
    info = fb.parse_cnf(cnf)
    atoms = np.ascontiguousarray(info["atoms"], np.float32)
    monkeypatch.chdir(tmp_path)
    out = {}
    devices = [0, 1] if torch.cuda.device_count() >= 2 else [0]
    for dev in devices:
        img = np.zeros((info["n3"], info["n2"], info["n1"]), np.float32)
        fb.cuda_FDES(dev, 0, str(cnf), str(tmp_path / f"m{dev}.bin"), str(tmp_path / f"r{dev}.emd"), atoms, len(atoms), img)
        out[dev] = img
    if len(devices) == 2:
        np.testing.assert_array_equal(out[0], out[1])
    monkeypatch.setenv("FDES_B200_GPUS", "2" if len(devices) == 2 else "0,0")
    img = np.zeros_like(out[0])
    fb.cuda_FDES(0, 0, str(cnf), str(tmp_path / "ms.bin"), str(tmp_path / "rs.emd"), atoms, len(atoms), img)
    assert rel_l2(img, out[0]) < 1e-6
    np.testing.assert_array_equal(np.fromfile(tmp_path / "ms.bin", np.float32).reshape(img.shape), img)
