"""Multi-GPU host logic on CPU ranks (gloo, world_size 2, 4, 6 and 8): slab
partition, halo exchange, all-gather layout and the global z-line tables, with
the kernels replaced by their numpy transcriptions (tests/emul.py).  The slab
results must equal the oracle / the single-domain run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import adi_oracle
import emul
import problems
import util


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, name, kwargs, nsteps, q, env=None):
    os.environ.update(env or {})
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import heatsim2_b200 as hs
        from heatsim2_b200 import dist as hdist
        prob = problems.ALL[name](hs, **kwargs)
        P, S = hdist.setup(*prob["setup_args"], plan_class=emul.EmulDistPlan)
        k0, k1 = P.slab
        T = torch.from_numpy(np.array(prob["T0"][k0:k1]))
        ve = prob["volumetric_elements"][k0:k1]
        for it in range(nsteps):
            T = hs.run_adi_steps(P, S, prob["t0"] + it * prob["dt"], prob["dt"], T, ve, prob["volumetric"])
        q.put((rank, k0, k1, T.numpy(), P.plan.comm_bytes_per_step()["interface_mode"], len(P.plan._line_ranges(T.shape[1] * T.shape[2]))))
    finally:
        dist.destroy_process_group()


def _run(world, name, kwargs, nsteps, env=None, info=None):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, kwargs, nsteps, q, env)) for r in range(world)]
    for p in procs:
        p.start()
    parts = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    parts.sort(key=lambda t: t[0])
    if info is not None:
        info["mode"] = parts[0][4]
        info["ranges"] = parts[0][5]
    return np.concatenate([p[3] for p in parts], axis=0)


@pytest.mark.parametrize("world,name,kwargs", [
    (2, "steelonfoam", dict(nz=32, ny=12, nx=14)),
    (4, "steelonfoam", dict(nz=64, ny=10, nx=12)),
    (2, "composite", dict(nz=32, ny=12, nx=16, ply=4)),
    (2, "steelonwater", dict(nz=48, ny=20, nx=24)),
    # all four source kinds; the z-decaying impulse carries a GLOBAL z_ndgrid
    (2, "sources_demo", dict(nz=16, ny=10, nx=14)),
    # BASELINE configs[4] recipe (uniform steel slab, random T0) on 8 ranks: strongly
    # implicit z-lines, so the interface band spans several 8-plane slabs
    (8, "uniform_slab", dict(shape=(64, 10, 12))),
])
def test_slabs_match_oracle(world, name, kwargs):
    import heatsim2_b200 as hs
    nsteps = 6 if name == "sources_demo" else 4
    got = _run(world, name, kwargs, nsteps)
    prob = problems.ALL[name](hs, **kwargs)
    want = adi_oracle.run(prob, nsteps=nsteps)
    assert util.relerr(got, want) <= 1e-12


def _worker_n(rank, world, port, name, kwargs, nsteps, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import heatsim2_b200 as hs
        from heatsim2_b200 import dist as hdist
        prob = problems.ALL[name](hs, **kwargs)
        P, S = hdist.setup(*prob["setup_args"], plan_class=emul.EmulDistPlan)
        k0, k1 = P.slab
        T0 = torch.from_numpy(np.array(prob["T0"][k0:k1]))
        ve = prob["volumetric_elements"][k0:k1]
        Tn, rec = hs.run_adi_steps_n(P, S, prob["t0"], prob["dt"], T0, ve, prob["volumetric"], nsteps,
                                     probes=[(1, 2, 3)], surface_dz=prob["dz"], every=2)
        q.put((rank, k0, k1, Tn.numpy(), rec))
    finally:
        dist.destroy_process_group()


def test_run_adi_steps_n_on_slab_plans():
    """run_adi_steps_n on a multi-rank plan (no native loop there: one library call per step, observation by tensor
    indexing on the rank's slab): field, probe and surface records against the oracle"""
    import heatsim2_b200 as hs
    kwargs, nsteps, world = dict(nz=32, ny=12, nx=14), 6, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_n, args=(r, world, port, "steelonfoam", kwargs, nsteps, q)) for r in range(world)]
    for p in procs:
        p.start()
    parts = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    prob = problems.steelonfoam(hs, **kwargs)
    O = adi_oracle.setup(*prob["setup_args"])
    T, hist = np.array(prob["T0"]), []
    for it in range(nsteps):
        T = O.step(prob["t0"] + it * prob["dt"], prob["dt"], T)
        if (it + 1) % 2 == 0:
            hist.append(T.copy())
    scale = np.abs(T).max()
    assert util.relerr(np.concatenate([p[3] for p in parts], axis=0), T) <= 1e-12
    for rank, k0, k1, _, rec in parts:
        assert rec["step"] == [2, 4, 6]
        want_probe = np.array([h[k0 + 1, 2, 3] for h in hist])
        assert np.abs(rec["probes"][:, 0] - want_probe).max() <= 1e-12 * scale
        # the surface estimate reads the slab's own first two planes (the physical z-min surface on rank 0)
        want_surf = np.array([hs.surface_temperature.insulating_z_min_surface_temperature(h[k0:k1], prob["dz"]) for h in hist])
        assert np.abs(rec["surface"] - want_surf).max() <= 1e-12 * scale


def test_neighbour_exchange_and_pipelining():
    """weakly coupled z-lines (thin plies): the interface band is a few chunks, so
    slabs exchange with their nearest neighbours only, in two pipelined line ranges"""
    import heatsim2_b200 as hs
    kwargs = dict(nz=96, ny=12, nx=16, ply=8)
    info = {}
    got = _run(6, "composite", kwargs, 3, env={"HS2_SLAB_CHUNK": "8", "HS2_DIST_MIN_LINES": "64"}, info=info)
    assert info["mode"].endswith(("1-hop neighbours", "2-hop neighbours")) and info["ranges"] == 2
    want = adi_oracle.run(problems.composite(hs, **kwargs), nsteps=3)
    assert util.relerr(got, want) <= 1e-12


def test_emulation_matches_oracle_single_domain():
    """the transcription itself (and hence the tables) is right"""
    import heatsim2_b200 as hs
    prob = problems.steelonfoam(hs, nz=24, ny=12, nx=14)
    P, S = hs.setup(*prob["setup_args"])
    O = adi_oracle.setup(*prob["setup_args"])
    T = np.array(prob["T0"])
    for it in range(3):
        src = O.volumetric_array(it * prob["dt"], prob["dt"])
        want = O.step(it * prob["dt"], prob["dt"], T)
        got = emul.step_single(P.plan, T, src)
        assert util.relerr(got, want) <= 1e-12
        T = want


def test_slab_partition_rules():
    from heatsim2_b200 import dist as hdist
    from heatsim2_b200.plan import slab_chunk
    assert hdist.slab_range(1024, 3, 8) == (384, 512)
    with pytest.raises(ValueError):
        hdist.slab_range(100, 0, 8)
    assert slab_chunk(128) == 32 and slab_chunk(16) == 16 and slab_chunk(24) == 8
    with pytest.raises(NotImplementedError):
        slab_chunk(20)
