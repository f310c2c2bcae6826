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
