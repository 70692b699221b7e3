"""N>1 host logic on CPU (gloo, world_size 2): slab partition maps + halo exchange ordering.

The numpy oracle computes the face traces of the global mesh; each rank then fetches exterior
traces through its LOCAL mapP with the halo filled by SlabHalo.exchange and must obtain exactly
what the global mapP gather gives.  This covers the send lists, the halo slot layout and the
two-messages-between-the-same-pair ordering of a 2-rank periodic ring.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, Kxyz, N, periodic_z, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from esdg_cns_b200 import halo as H, setup_dg as SD
        prd, _ = SD.reference_hex(N)
        per = (True, True, periodic_z)
        mdg = SD.uniform_hex_mesh_md(prd, Kxyz, periodic=per)
        mdl = SD.uniform_hex_mesh_md(prd, Kxyz, periodic=per, zslab=(rank, world))
        Nfq = 6 * (N + 1) ** 2
        nf = 3
        rng = np.random.default_rng(7)
        Tg = rng.random((nf, Nfq * mdg.K))                    # a global "trace" array, field-major
        k0, k1 = mdl.zrange
        e0, e1 = k0 * Kxyz[0] * Kxyz[1], k1 * Kxyz[0] * Kxyz[1]
        Tl = np.zeros((nf, Nfq * mdl.K + mdl.n_halo))
        Tl[:, :Nfq * mdl.K] = Tg[:, e0 * Nfq:e1 * Nfq]
        h = H.SlabHalo(mdl, nf, dist, torch, "cpu")
        h.send.copy_(torch.from_numpy(Tl[:, h.send_list]))
        recv = h.exchange().numpy()
        Tl[:, Nfq * mdl.K:] = recv
        mine = Tl[:, mdl.mapP.flatten(order="F")]
        want = Tg[:, mdg.mapP[:, e0:e1].flatten(order="F")]
        # non-periodic outer boundary: local map points to itself like the global one
        q.put((rank, float(np.abs(mine - want).max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,periodic_z", [(2, True), (2, False), (3, True)])
def test_slab_halo_matches_global_gather(world, periodic_z):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world * 2 + int(periodic_z)
    procs = [ctx.Process(target=_worker, args=(r, world, port, (3, 2, 6), 2, periodic_z, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world))
    assert all(v == 0.0 for v in res.values()), res


def _worker_quad(rank, world, port, Kxy, N, periodic_y, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from esdg_cns_b200 import halo as H, setup_dg as SD
        prd, _ = SD.reference_quad_modal(N)
        per = (False, periodic_y)
        mdg = SD.uniform_quad_mesh_md(prd, *Kxy, periodic=per)
        mdl = SD.uniform_quad_mesh_md(prd, *Kxy, periodic=per, yslab=(rank, world))
        Nfq = 4 * (N + 1)
        rng = np.random.default_rng(11)
        j0, j1 = mdl.yrange
        e0, e1 = j0 * Kxy[0], j1 * Kxy[0]
        worst = 0.0
        for nf in (4, 6):                                      # entropy-variable traces, sigma traces (CNS: two exchanges)
            Tg = rng.random((nf, Nfq * mdg.K))
            Tl = np.zeros((nf, Nfq * mdl.K + mdl.n_halo))
            Tl[:, :Nfq * mdl.K] = Tg[:, e0 * Nfq:e1 * Nfq]
            h = H.SlabHalo(mdl, nf, dist, torch, "cpu")
            h.send.copy_(torch.from_numpy(Tl[:, h.send_list]))
            Tl[:, Nfq * mdl.K:] = h.exchange().numpy()
            mine = Tl[:, mdl.mapP.flatten(order="F")]
            want = Tg[:, mdg.mapP[:, e0:e1].flatten(order="F")]
            worst = max(worst, float(np.abs(mine - want).max()))
        # physical boundary nodes of the slab = the global ones in its element range
        mapB_l = np.nonzero(mdl.mapP.flatten(order="F") == np.arange(Nfq * mdl.K))[0]
        gB = np.nonzero(mdg.mapP.flatten(order="F") == np.arange(Nfq * mdg.K))[0]
        gB = gB[(gB >= e0 * Nfq) & (gB < e1 * Nfq)] - e0 * Nfq
        q.put((rank, worst, bool(np.array_equal(mapB_l, gB))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,periodic_y", [(2, False), (2, True), (3, False)])
def test_quad_yslab_halo_matches_global_gather(world, periodic_y):
    """y-slab partition of the quad mesh used by the distributed CNS path (two exchanges per stage)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + world * 2 + int(periodic_y)
    procs = [ctx.Process(target=_worker_quad, args=(r, world, port, (4, 6), 2, periodic_y, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res = [q.get(timeout=5) for _ in range(world)]
    assert all(r[1] == 0.0 and r[2] for r in res), res


def test_exchange_in_process_equals_gloo_layout():
    """The single-process exchange (device copies instead of send/recv) fills the halo like the gloo one."""
    from esdg_cns_b200 import halo as H, setup_dg as SD
    N, Kxy, world = 2, (3, 6), 3
    prd, _ = SD.reference_quad_modal(N)
    mdg = SD.uniform_quad_mesh_md(prd, *Kxy, periodic=(False, False))
    Nfq = 4 * (N + 1)
    Tg = np.random.default_rng(5).random((4, Nfq * mdg.K))
    mds = [SD.uniform_quad_mesh_md(prd, *Kxy, periodic=(False, False), yslab=(r, world)) for r in range(world)]
    hs = [H.SlabHalo(md, 4, None, torch, "cpu") for md in mds]
    for md, h in zip(mds, hs):
        e0 = md.yrange[0] * Kxy[0]
        h.send.copy_(torch.from_numpy(Tg[:, e0 * Nfq + h.send_list]))
    H.exchange_in_process(hs)
    for md, h in zip(mds, hs):
        e0, e1 = md.yrange[0] * Kxy[0], md.yrange[1] * Kxy[0]
        Tl = np.concatenate([Tg[:, e0 * Nfq:e1 * Nfq], h.recv.numpy()], axis=1)
        assert np.array_equal(Tl[:, md.mapP.flatten(order="F")], Tg[:, mdg.mapP[:, e0:e1].flatten(order="F")])
