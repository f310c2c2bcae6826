"""Multi-GPU driver: one process per GPU, frozen-phonon configurations sharded across ranks.

The reference is single-GPU (no NCCL/MPI anywhere); its driver loop
(src/crystalMaker.cu:324-373) averages the configurations j of every measurement k into the
intensity I_k and the coherent exit wave E_k and only then applies the detector tail once
(addNoiseAndMtf, :372).  That structure shards naturally:

    rank r runs configurations [count*r/world, count*(r+1)/world) of every k   (no data exchange)
    all_reduce(SUM) of the partial I_k (m1*m2 float32) and, if wanted, E_k     (the only collective)
    detector tail on the reduced I_k                                           (every rank, identical)

Every rank seeds the XORWOW streams like the reference (curand_init(1, i, 0)) and discards the
normals of the configurations before its own, so configuration j sees the same displacements as
in a single-GPU run.  torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) is
plumbing only: the partial sums are written by the library's kernels straight into the tensors
that are reduced.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def shard_range(count: int, rank: int, world: int) -> Tuple[int, int]:
    """Configurations [begin, end) of `rank`; contiguous blocks, sizes differ by at most one.
    Same arithmetic as Engine::Engine in fdes_b200/csrc/engine.cu."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return (count * rank) // world, (count * (rank + 1)) // world


def simulate_sharded(open_sim: Callable[[int, int], object], *, want_exitwave: bool = False, group=None,
                     device=None, timings: Optional[dict] = None):
    """Run a whole simulation with the configurations sharded over the ranks of `group`.

    open_sim(rank, world) must return an object with the session interface of
    fdes_b200.Simulation (n1 n2 n3 m1 m2, set_accumulators, run_k, finish_k, close) that handles
    this rank's share.  Returns (image [n3, n2, n1] float32, exitwave [n3, m2, m1] complex64 | None)
    -- identical on every rank.  `timings` (optional dict) receives "collective_ms": the time this
    rank spent in the all-reduces (CUDA events on GPUs, wall clock on CPU).
    """
    import time
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    sim = open_sim(rank, world)
    try:
        # the accumulators must live on the device the session computes on (not torch's current one)
        if device is not None:
            dev = torch.device(device)
        elif torch.cuda.is_available() and getattr(sim, "gpu_index", None) is not None:
            dev = torch.device("cuda", int(sim.gpu_index))
        else:
            dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        acc_I = torch.zeros(sim.m2 * sim.m1, dtype=torch.float32, device=dev)
        acc_E = torch.zeros(sim.m2 * sim.m1 * 2, dtype=torch.float32, device=dev) if want_exitwave else None
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)        # torch's fill runs on torch's stream, the engine on its own
        sim.set_accumulators(acc_I.data_ptr(), acc_E.data_ptr() if acc_E is not None else 0)
        coll_ms = 0.0
        image = np.zeros((sim.n3, sim.n2, sim.n1), np.float32)
        exitwave = np.zeros((sim.n3, sim.m2, sim.m1), np.complex64) if want_exitwave else None
        for k in range(sim.n3):
            sim.run_k(k)                       # partial sums of this rank's configurations
            if world > 1:
                if dev.type == "cuda":
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    with torch.cuda.device(dev):
                        e0.record()
                else:
                    t0 = time.perf_counter()
                dist.all_reduce(acc_I, op=dist.ReduceOp.SUM, group=group)
                if acc_E is not None:
                    dist.all_reduce(acc_E, op=dist.ReduceOp.SUM, group=group)
                if dev.type == "cuda":
                    with torch.cuda.device(dev):
                        e1.record()
                    torch.cuda.synchronize(dev)
                    coll_ms += e0.elapsed_time(e1)
                else:
                    coll_ms += (time.perf_counter() - t0) * 1e3
            img, ew = sim.finish_k(k)          # detector tail on the reduced intensity
            image[k] = img
            if exitwave is not None:
                exitwave[k] = ew
        if timings is not None:
            timings["collective_ms"] = coll_ms
        return image, exitwave
    finally:
        sim.close()


def stem_scan_sharded(open_sim: Callable[[int, int], object], positions, detectors_mrad, *, k: int = 0, group=None,
                      device=None):
    """STEM scan with the probe positions sharded over the ranks of `group` (contiguous blocks of the
    raster, the same ranges as `shard_range`); every rank builds the transmission stack of each
    frozen-phonon configuration itself (no exchange), and one all-gather of the detector signals
    ([n_probes, n_detectors] float32, <= 512 KB for a 256 x 256 scan) gives every rank the full result.

    open_sim(rank, world) returns an object with `stem_scan(positions, detectors_mrad, k) -> (signals, ms)`
    (fdes_b200.Simulation opened WITHOUT rank/world sharding: all configurations on every rank)."""
    import torch
    import torch.distributed as dist

    positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 2)
    det = np.ascontiguousarray(detectors_mrad, np.float32).reshape(-1, 2)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(len(positions), rank, world)
    sim = open_sim(rank, world)
    try:
        mine, _ = sim.stem_scan(positions[lo:hi], det, k) if hi > lo else (np.zeros((0, len(det)), np.float32), 0.0)
    finally:
        sim.close()
    if world == 1:
        return mine
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                             if torch.cuda.is_available() else torch.device("cpu"))
    longest = max(b - a for a, b in (shard_range(len(positions), r, world) for r in range(world)))
    send = torch.zeros((longest, len(det)), dtype=torch.float32, device=dev)
    send[: hi - lo] = torch.from_numpy(np.ascontiguousarray(mine)).to(dev)
    parts = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(parts, send, group=group)
    out = [parts[r][: b - a].cpu().numpy() for r, (a, b) in
           ((r, shard_range(len(positions), r, world)) for r in range(world))]
    return np.concatenate(out, axis=0)
