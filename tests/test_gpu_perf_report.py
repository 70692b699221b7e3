"""Throughput report of the 2-D configurations of BASELINE.json that bench.py does not time (config 2: modal
triangles N=4, 256 x 256 x 2 elements; config 1 / 3 shapes at scale are in profiles/measure_2d.py).  The triangle
reference element comes from the oracle's setup (the package has structured quad / hex setup only), so this lives
under tests/; the big mesh is the oracle's 8 x 8 periodic mesh replicated as independent periodic tiles, which keeps
the per-element work and the operator tables of config 2 exactly and only changes which tile a neighbour lives in."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_tri_config2_throughput_report():
    from types import SimpleNamespace
    from oracle import solvers as SV
    from esdg_cns_b200 import api
    N, tiles = 4, 1024                                   # 128 triangles per tile -> K = 131 072 (config 2)
    rd, md, ops, Q = SV.setup_tri_vortex(N, 8, 8)
    K0, Nfq = md.K, md.mapP.shape[0]
    big = SimpleNamespace(K=K0 * tiles, dim=2)
    for nme in ("rxJ", "sxJ", "ryJ", "syJ", "J", "nxJ", "nyJ", "sJ"):
        setattr(big, nme, np.tile(np.atleast_2d(getattr(md, nme)), (1, tiles)))
    off = (np.arange(tiles) * (Nfq * K0))[None, :, None]
    big.mapP = (md.mapP[:, None, :] + off).reshape(Nfq, K0 * tiles, order="C") if False else \
        np.concatenate([md.mapP + t * Nfq * K0 for t in range(tiles)], axis=1)
    Qb = [np.tile(q, (1, tiles)) for q in Q]
    s = api.Solver(big, ops)
    # the tiled run reproduces the small one (same elements, same neighbours)
    ref = api.Solver(md, ops)
    r_small, _ = ref.rhs(Q, False)
    r_big, _ = s.rhs(Qb, False)
    assert all(np.array_equal(a, b[:, :K0]) and np.array_equal(a, b[:, -K0:]) for a, b in zip(r_small, r_big))
    s.upload(Qb)
    dt = 1e-4
    s.lsrk45(dt, 2)
    s.sync()
    s.enable_timing(True)
    t0 = time.perf_counter()
    nsteps = 10
    s.lsrk45(dt, nsteps)
    s.sync()
    wall = time.perf_counter() - t0
    dof = Qb[0].shape[0] * big.K * 4
    ms_rhs, nl = s.kernel_time_ms("generic_rhs")
    ms_tr, _ = s.kernel_time_ms("generic_traces")
    print(f"tri N={N} K={big.K} (config 2): {wall / nsteps / 5 * 1e3:.3f} ms/stage, {dof * 5 * nsteps / wall * 1e-9:.2f} GDOF*stage/s; "
          f"generic_rhs {ms_rhs / nl:.3f} ms, generic_traces {ms_tr / nl:.3f} ms per launch")
    assert all(np.isfinite(q).all() for q in s.download())
