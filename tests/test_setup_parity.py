"""esdg_cns_b200.setup_dg (tensor-product, analytic connectivity) must reproduce the oracle's literal
restatement of SetupDG.jl / connect_mesh / build_node_maps / geometric_factors array by array."""
import numpy as np
import pytest

from esdg_cns_b200 import setup_dg as SD
from oracle import solvers as SV


def _close(a, b, tol=1e-13):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


@pytest.mark.parametrize("N,Kx,Ky", [(1, 3, 2), (3, 5, 4)])
def test_quad_setup_matches_oracle(N, Kx, Ky):
    rd, md, ops, _ = SV.setup_quad_vortex(N, Kx, Ky)
    prd, pops = SD.reference_quad(N)
    pmd = SD.uniform_quad_mesh_md(prd, Kx, Ky, domain=((0, 15), (-5, 5)))
    for n in ("rq", "sq", "wq", "wf", "rf", "sf", "nrJ", "nsJ"):
        assert _close(getattr(prd, n), getattr(rd, n)), n
    assert _close(pops.Ef, ops.Ef) and _close(pops.Ph, ops.Ph) and _close(pops.Lf, ops.Lf)
    for a, b in zip(pops.Qskew, ops.Qskew):
        assert _close(a, b)
    assert _close(pmd.xq, md.xq) and _close(pmd.yq, md.yq)
    nfp = N + 1
    for n in ("rxJ", "sxJ", "ryJ", "syJ", "J"):
        assert _close(np.broadcast_to(getattr(pmd, n), getattr(md, n).shape), getattr(md, n)), n
    for n in ("nxJ", "nyJ", "sJ"):
        assert _close(np.repeat(getattr(pmd, n), nfp, axis=0), getattr(md, n)), n
    assert np.array_equal(pmd.mapP, md.mapP)


@pytest.mark.parametrize("N,K", [(1, (2, 3, 2)), (2, (3, 2, 4)), (3, (2, 2, 2))])
def test_hex_setup_matches_oracle(N, K):
    rd, md, ops = SV.setup_hex(N, K, scale=(1.0, 2.0, .5), shift=(.1, .2, .3))
    prd, pops = SD.reference_hex(N)
    pmd = SD.uniform_hex_mesh_md(prd, K, domain=((-.9, 1.1), (-1.8, 2.2), (-.2, .8)))
    for n in ("rq", "sq", "tq", "wq", "wf", "rf", "sf", "tf", "nrJ", "nsJ", "ntJ"):
        assert _close(getattr(prd, n), getattr(rd, n)), n
    assert _close(pops.Ef, ops.Ef) and _close(pops.Ph, ops.Ph) and _close(pops.Lf, ops.Lf)
    for a, b in zip(pops.Qskew, ops.Qskew):
        assert _close(a, b)
    x, y, z = SD.hex_node_coords(prd, pmd)
    assert _close(x, md.xq) and _close(y, md.yq) and _close(z, md.zq)
    for n in ("rxJ", "sxJ", "txJ", "ryJ", "syJ", "tyJ", "rzJ", "szJ", "tzJ", "J"):
        assert _close(np.broadcast_to(getattr(pmd, n), getattr(md, n).shape), getattr(md, n)), n
    nfp = (N + 1) ** 2
    for n in ("nxJ", "nyJ", "nzJ", "sJ"):
        assert _close(np.repeat(getattr(pmd, n), nfp, axis=0), getattr(md, n)), n
    assert np.array_equal(pmd.mapP, md.mapP)


def test_nonperiodic_boundary_maps_to_itself():
    prd, _ = SD.reference_hex(1)
    md = SD.uniform_hex_mesh_md(prd, (2, 2, 2), periodic=(False, True, True))
    self_idx = np.arange(md.mapP.size).reshape(md.mapP.shape, order="F")
    assert (md.mapP == self_idx).sum() == 2 * 4 * 4    # two x-boundaries, 4 faces each, 4 points per face


def test_modal_quad_and_cavity_sets_match_oracle():
    """reference_quad_modal (tensor-product restatement of init_reference_quad + the CNS operator
    tuple) and the cavity boundary sets against the oracle's literal SetupDG restatement."""
    from oracle import cns
    for N in (2, 3):
        prd, pops = SD.reference_quad_modal(N)
        rd, md, ops, bc, _ = cns.setup_cavity("quad", N, 3)
        for n in ("r", "s", "rq", "sq", "wq", "rf", "sf", "wf", "Dr", "Ds", "Vq", "Pq", "Vf", "LIFT"):
            assert np.abs(getattr(prd, n) - getattr(rd, n)).max() < 1e-12, n
        for n in ("Ph", "VhP", "Lf"):
            assert np.abs(getattr(pops, n) - getattr(ops, n)).max() < 1e-11, n
        for a, b in zip(pops.Qskew, ops.Qskew):
            assert np.abs(a - b).max() < 1e-13
        pmd = SD.uniform_quad_mesh_md(prd, 3, 3, periodic=(False, False))
        assert np.array_equal(pmd.mapP, md.mapP)
        assert np.abs(pmd.x - md.x).max() < 1e-14 and np.abs(pmd.yf - md.yf).max() < 1e-14
        pbc = SD.cavity_bc_sets(pmd)
        assert np.array_equal(np.sort(pbc.lid), np.sort(bc.lid))
        assert np.array_equal(np.sort(pbc.wall), np.sort(bc.wall))
        assert np.array_equal(np.sort(pbc.mapB), np.sort(md.mapB))


def test_curved_hex_mesh_of_the_package_is_watertight_and_preserves_the_free_stream():
    """setup_dg.curved_hex_mesh_md (displaced vertices, trilinear elements, curl-form geometric factors) driven through the
    ORACLE's hex rhs: a constant state has zero rhs (metric identities + matching normals on shared faces), the rhs without
    penalty conserves entropy to round-off, and the mesh really is curved."""
    from oracle import solvers as SV, physics as P
    from esdg_cns_b200 import setup_dg as SD
    N = 2
    rd, md0, ops = SV.setup_hex(N, 3)
    prd, pops = SD.reference_hex(N)
    md = SD.curved_hex_mesh_md(prd, (3, 4, 3), domain=((-1, 1), (-1.5, 1.5), (0, 2)), amp=.1)
    md.wJq = np.diag(rd.wq) @ md.J
    assert md.rxJ.shape == (ops.Ef.shape[0] + rd.wq.size, md.K) and md.J.min() > 0
    assert np.abs(md.rxJ - md.rxJ[:1]).max() > 1e-3 * np.abs(md.rxJ).max()
    # normals of the two sides of every face are opposite (watertight, conforming)
    for a in (md.nxJ, md.nyJ, md.nzJ):
        assert np.abs(a + a.reshape(-1, order="F")[md.mapP.reshape(-1, order="F")].reshape(a.shape, order="F")).max() < 1e-13
    one = np.ones_like(md.xq)
    Qc = list(P.primitive_to_conservative(1.2 * one, (.3 * one, -.2 * one, .1 * one), .9 * one))
    out, _ = SV.rhs_collocated(Qc, md, ops, rd, .25, False)
    assert max(np.abs(x).max() for x in out) < 1e-12
    rng = np.random.default_rng(0)
    rho, p = 2 + .1 * rng.random(one.shape), 1 + .1 * rng.random(one.shape)
    Q = list(P.primitive_to_conservative(rho, (.3 + .1 * rng.random(one.shape), 1 + .1 * rng.random(one.shape), -.2 + 0 * one), p))
    _, rt = SV.rhs_collocated(Q, md, ops, rd, 0.0, True)
    assert abs(rt) < 1e-11


@pytest.mark.parametrize("N,Kx,Ky", [(1, 3, 2), (2, 2, 3), (3, 4, 3), (4, 3, 3)])
def test_tri_setup_matches_oracle(N, Kx, Ky):
    """setup_dg.reference_tri / uniform_tri_mesh_md (own basis formulation, analytic connectivity) against the oracle's
    literal restatement of init_reference_tri + the generic mesh pipeline (connect_mesh, build_node_maps, periodic maps)."""
    rd, md, ops, _ = SV.setup_tri_vortex(N, Kx, Ky)
    prd, pops = SD.reference_tri(N)
    pmd = SD.uniform_tri_mesh_md(prd, Kx, Ky, domain=((0, 15), (-5, 5)))
    for nm in ("r", "s", "rq", "sq", "wq", "rf", "sf", "wf", "nrJ", "nsJ", "Dr", "Ds", "Vq", "Pq", "Vf", "LIFT"):
        assert np.abs(np.asarray(getattr(rd, nm)) - np.asarray(getattr(prd, nm))).max() < 1e-11, nm
    for nm in ("Ph", "VhP", "Lf", "Vq"):
        assert np.abs(getattr(ops, nm) - getattr(pops, nm)).max() < 1e-11, nm
    for a, b in zip(ops.Qskew, pops.Qskew):
        assert np.abs(a - b).max() < 1e-11
    assert pops.volume_scale == ops.volume_scale and pops.lf_scale == ops.lf_scale
    n = N + 1
    for nm in ("x", "y", "xq", "yq", "xf", "yf"):
        assert np.abs(getattr(md, nm) - getattr(pmd, nm)).max() < 1e-12, nm
    for nm in ("rxJ", "sxJ", "ryJ", "syJ"):
        assert np.abs(getattr(md, nm)[:1] - getattr(pmd, nm)).max() < 1e-12, nm
    for nm in ("nxJ", "nyJ", "sJ"):
        assert np.abs(getattr(md, nm) - np.repeat(getattr(pmd, nm), n, axis=0)).max() < 1e-12, nm
    assert np.abs(md.J[:1] - pmd.J).max() < 1e-12
    assert np.array_equal(md.mapP, pmd.mapP)
