"""z-slab multi-GPU path on real GPUs (NCCL): N ranks must reproduce the
single-GPU field (<= 1e-13 relative) and the oracle."""
import os
import socket

import numpy as np
import pytest

import problems
import util

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, name, kwargs, nsteps, q, p2p, same_gpu=False, fused=True):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["HS2_DIST_P2P"] = "1" if p2p else "0"
    os.environ["HS2_DIST_TIMEOUT_S"] = "30" if same_gpu else "10"
    os.environ["HS2_DIST_Z_FUSED"] = "1" if fused else "0"
    os.environ["HS2_DIST_MIN_LINES"] = "512"          # two line ranges in the peer-memory z sweep on these small grids
    if same_gpu:
        # every rank on GPU 0: the peer-memory transport (CUDA IPC mailboxes, flags, stores from the z kernel)
        # works between processes of one device too; only the set-up plumbing needs a backend (gloo)
        torch.cuda.set_device(0)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import heatsim2_b200 as hs
        from heatsim2_b200 import dist as hdist
        prob = problems.ALL[name](hs, **kwargs)
        P, S = hdist.setup(*prob["setup_args"])
        k0, k1 = P.slab
        T = torch.from_numpy(np.array(prob["T0"][k0:k1])).cuda()
        ve = prob["volumetric_elements"][k0:k1]
        for it in range(nsteps):
            T = hs.run_adi_steps(P, S, prob["t0"] + it * prob["dt"], prob["dt"], T, ve, prob["volumetric"])
        torch.cuda.synchronize()
        assert (P.plan._px is not None) == bool(p2p)
        P.plan.check()
        q.put((rank, T.cpu().numpy()))
        P.plan.close()
    except Exception as exc:          # report instead of leaving the parent waiting
        import traceback
        q.put((rank, "ERROR in rank %d: %s\n%s" % (rank, exc, traceback.format_exc())))
        raise
    finally:
        dist.destroy_process_group()


def _run(world, name, kwargs, nsteps, p2p=True, same_gpu=False, fused=True):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, kwargs, nsteps, q, p2p, same_gpu, fused)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        parts = [q.get(timeout=240) for _ in range(world)]
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
    for part in parts:
        assert not isinstance(part[1], str), part[1]
    for p in procs:
        assert p.exitcode == 0
    parts.sort(key=lambda t: t[0])
    return np.concatenate([p[1] for p in parts], axis=0)


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("p2p", [True, "ranges", False], ids=["peer-memory-fused", "peer-memory-ranges", "nccl"])
@pytest.mark.parametrize("name,kwargs,nsteps", [
    ("steelonfoam", dict(nz=64, ny=40, nx=48), 6),
    ("uniform_slab", dict(shape=(128, 48, 64)), 4),
    ("composite", dict(nz=64, ny=32, nx=32, ply=8), 4),
    ("sources_demo", dict(nz=16, ny=10, nx=14), 6),
    ("steelonwater", dict(nz=64, ny=24, nx=512), 3),        # warp x kernel in slabs (interior / boundary planes)
])
def test_two_gpus_match_one_gpu_and_oracle(name, kwargs, nsteps, p2p):
    import adi_oracle
    import heatsim2_b200 as hs
    world = min(_ngpu(), 4) if name == "uniform_slab" else 2
    got = _run(world, name, kwargs, nsteps, bool(p2p), fused=p2p is True)
    prob = problems.ALL[name](hs, **kwargs)
    one = util.run_b200(hs, prob, nsteps=nsteps)
    assert util.relerr(got, one) <= 1e-13
    assert util.relerr(got, adi_oracle.run(prob, nsteps=nsteps)) <= 1e-12


@pytest.mark.skipif(_ngpu() < 1, reason="needs a GPU")
@pytest.mark.parametrize("world,name,kwargs,nsteps,fused", [
    (2, "steelonfoam", dict(nz=64, ny=40, nx=48), 4, True),
    (2, "steelonfoam", dict(nz=64, ny=40, nx=48), 4, False),
    (4, "uniform_slab", dict(shape=(128, 48, 64)), 3, True),
    (4, "uniform_slab", dict(shape=(128, 48, 64)), 3, False),
    (2, "sources_demo", dict(nz=16, ny=10, nx=14), 6, True),
    (2, "composite", dict(nz=64, ny=96, nx=128, ply=8), 3, True),      # 768 tiles: several per block, two line classes
    (2, "uniform_slab", dict(shape=(256, 24, 40)), 3, True),           # 4 local chunks of 32: warp-autonomous fused kernel, 8 lines per warp
    (2, "composite", dict(nz=128, ny=24, nx=40, ply=8), 3, True),      # 2 local chunks: 16 lines per warp, two line classes
    (2, "uniform_slab", dict(shape=(512, 16, 24)), 3, True),           # 8 local chunks: block kernel, two line groups with their own barriers
])
def test_peer_memory_transport_between_processes_on_one_gpu(world, name, kwargs, nsteps, fused):
    """The mailbox / flag / step-parity protocol of the peer-memory transport (dist.py PeerExchange, peer.cu,
    z_forward's stores into the peers' rows) with every rank a process on GPU 0 - runs on a one-GPU box.
    The GPU time-slices between the processes, so a flag wait costs a time slice; a few steps are enough."""
    import adi_oracle
    import heatsim2_b200 as hs
    got = _run(world, name, kwargs, nsteps, p2p=True, same_gpu=True, fused=fused)
    prob = problems.ALL[name](hs, **kwargs)
    one = util.run_b200(hs, prob, nsteps=nsteps)
    assert util.relerr(got, one) <= 1e-13
    assert util.relerr(got, adi_oracle.run(prob, nsteps=nsteps)) <= 1e-12
