"""GPU parity of the tensor-product QUAD kernels (esdg_cns_b200/csrc/kernels_quad_line.cuh; BASELINE config 1,
dg2D_euler_quad.jl:141-207) against the numpy oracle, the generic table-driven kernels and the LSRK45 loop.

Tolerance: the north star's 1e-12 relative in FP64 (helpers.parity_tol raises it to 4x the oracle's own rounding cloud
where the reference's logmean formula is conditioned worse than that).
"""
import numpy as np
import pytest

from helpers import relerr as _relerr, parity_tol

pytestmark = pytest.mark.gpu


def _sheared(N, Kx, Ky):
    """The vortex mesh of config 1 under a linear shear: every element an (affine) parallelogram, all four geometric
    factors non-zero.  Connectivity (periodic mapP) is that of the unsheared mesh."""
    from oracle import solvers as SV, setupdg as S, mesh as M, physics as P
    rd, md0, ops, _ = SV.setup_quad_vortex(N, Kx, Ky)
    VX, VY, EToV = M.uniform_quad_mesh(Kx, Ky)
    VX = 15 * (1 + VX) / 2
    VY = 5 * VY
    md = S.init_mesh((VX + .3 * VY, VY + .1 * VX), EToV, rd)
    md.mapP = md0.mapP
    SV.interp_geofacs(md, rd, ops)
    rho, u, v, p = P.vortex(md0.xq, md0.yq, 0)
    Q = list(P.primitive_to_conservative(rho, (u, v), p))
    return rd, md, ops, Q


@pytest.mark.parametrize("N,Kx,Ky", [(1, 5, 4), (2, 7, 3), (3, 16, 16), (4, 6, 9)])
@pytest.mark.parametrize("lf", [0.0, 0.5])
def test_quad_line_rhs_matches_oracle(N, Kx, Ky, lf):
    from oracle import solvers as SV
    from esdg_cns_b200 import api
    rd, md, ops, Q = SV.setup_quad_vortex(N, Kx, Ky)
    ops.lf_scale = lf
    ref, rt_ref = SV.rhs_collocated(Q, md, ops, rd, lf, True)
    s = api.Solver(md, ops)
    assert s.kernel_path() == "quad_line" and s.trace_planes() == 7
    out, rt = s.rhs(Q, True)
    tol = parity_tol(lambda: SV.rhs_collocated(Q, md, ops, rd, lf, False)[0], ref)
    err = _relerr(out, ref)
    print(f"quad N={N} {Kx}x{Ky} lf={lf}: relerr {err:.2e} tol {tol:.2e} rhstest {rt:.3e} (oracle {rt_ref:.3e})")
    assert err < tol
    assert abs(rt - rt_ref) < 1e-11 * max(1.0, abs(rt_ref))
    if lf == 0.0:
        assert abs(rt) < 1e-11


@pytest.mark.parametrize("N", [2, 3])
def test_quad_line_sheared_mesh(N):
    """Parallelogram elements: the general (non-diagonal) metric contraction and oblique normals."""
    from oracle import solvers as SV
    from esdg_cns_b200 import api
    rd, md, ops, Q = _sheared(N, 6, 5)
    ref, rt_ref = SV.rhs_collocated(Q, md, ops, rd, .5, True)
    s = api.Solver(md, ops)
    assert s.kernel_path() == "quad_line"
    out, rt = s.rhs(Q, True)
    err = _relerr(out, ref)
    tol = parity_tol(lambda: SV.rhs_collocated(Q, md, ops, rd, .5, False)[0], ref)
    print(f"sheared quad N={N}: relerr {err:.2e} tol {tol:.2e}")
    assert err < tol
    assert abs(rt - rt_ref) < 1e-11 * max(1.0, abs(rt_ref))


def test_quad_line_matches_generic_kernels(monkeypatch):
    """Same operators through the table-driven kernels (ESDG_FORCE_GENERIC): rhs and 3 LSRK45 steps agree within the
    rounding cloud of the reference formula (both sit ~1e-12 from the oracle on the vortex's near-uniform far field)."""
    from oracle import solvers as SV
    from esdg_cns_b200 import api
    rd, md, ops, Q = SV.setup_quad_vortex(3, 9, 7)
    res = []
    for force in (False, True):
        if force:
            monkeypatch.setenv("ESDG_FORCE_GENERIC", "1")
        s = api.Solver(md, ops)
        assert s.kernel_path() == ("quad_line" if not force else s.kernel_path())
        r, rt = s.rhs(Q, True)
        s.upload(Q)
        rts = s.lsrk45(1e-2, 3, want_rhstest=True)
        res.append((r, rt, s.download(), rts, s.kernel_path()))
        s.close()
    assert res[0][4] == "quad_line" and res[1][4] != "quad_line"
    ref = SV.rhs_collocated(Q, md, ops, rd, .5, False)[0]
    assert _relerr(res[0][0], res[1][0]) < 2 * parity_tol(lambda: SV.rhs_collocated(Q, md, ops, rd, .5, False)[0], ref)
    assert abs(res[0][1] - res[1][1]) < 1e-12
    assert max(np.abs(a - b).max() for a, b in zip(res[0][2], res[1][2])) < 1e-12
    assert abs(res[0][3] - res[1][3]) < 1e-12


def test_quad_line_lsrk45_matches_oracle_steps():
    """Config 1 at the script's CFL: 5 LSRK45 steps (25 fused stage launches) against the oracle's loop."""
    from oracle import solvers as SV
    from esdg_cns_b200 import api
    N, K1D = 3, 16
    rd, md, ops, Q = SV.setup_quad_vortex(N, K1D, K1D)
    dt = 2 * (2 / K1D) / ((N + 1) * (N + 2) / 2)
    Qo, res = [q.copy() for q in Q], [np.zeros_like(q) for q in Q]
    for _ in range(5):
        Qo, res, rto = SV.lsrk45_step(Qo, res, dt, lambda q, t: SV.rhs_collocated(q, md, ops, rd, .5, t))
    s = api.Solver(md, ops)
    s.upload(Q)
    n0 = s.kernel_launches()
    rtg = s.lsrk45(dt, 5, want_rhstest=True)
    Qg = s.download()
    err = max(np.abs(a - b).max() for a, b in zip(Qg, Qo))
    print(f"quad config 1, 5 steps: max abs err {err:.2e}; launches {s.kernel_launches() - n0}")
    assert err < 5e-12          # 25 stages of 1e-13-level rhs differences on states of magnitude 1..10
    assert abs(rtg - rto) < 1e-12


def test_quad_line_ragged_last_batch_and_freestream():
    """K not a multiple of the CTA batch (16 elements at N=3); a constant state has zero rhs."""
    from esdg_cns_b200 import api, setup_dg as SD
    prd, pops = SD.reference_quad(3)
    pmd = SD.uniform_quad_mesh_md(prd, 7, 5)
    s = api.Solver(pmd, pops)
    assert s.kernel_path() == "quad_line"
    Q = [np.full((16, 35), v) for v in (1.1, .3, -.2, 2.5)]
    out, _ = s.rhs(Q, False)
    assert max(np.abs(r).max() for r in out) < 1e-12


def test_quad_line_domain_error():
    from esdg_cns_b200 import api, setup_dg as SD, _lib
    prd, pops = SD.reference_quad(2)
    pmd = SD.uniform_quad_mesh_md(prd, 4, 4)
    Q = [np.full((9, 16), v) for v in (1.0, .3, -.2, 2.5)]
    Q[0][4, 7] = -1.0
    with pytest.raises(_lib.DomainError):
        api.rhs(Q, pmd, pops)


@pytest.mark.parametrize("N,mesh,nslabs,overlap", [(3, (6, 8), 2, True), (2, (5, 9), 3, True), (3, (4, 6), 3, False), (4, (5, 3), 3, True)])
def test_quad_yslabs_with_halo_exchange_match_single_context(N, mesh, nslabs, overlap):
    """SURVEY 8e on the quad path: y-slabs with halo slots, boundary rows first / interior second, 7-plane pack and
    unpack -- the call sequence of the NCCL stepper with device-to-device copies as transport.  Bitwise equal to the
    single-context run."""
    import torch
    from esdg_cns_b200 import api, halo as H, setup_dg as SD
    prd, pops = SD.reference_quad(N)
    dom = ((0.0, 20.0), (-5.0, 5.0))
    mdg = SD.uniform_quad_mesh_md(prd, *mesh, domain=dom)
    rho, u, v, p = SD.vortex(mdg.xq, mdg.yq)
    Q = SD.primitive_to_conservative(rho, (u, v), p)
    dt, nsteps = 5e-3, 2
    stream = torch.cuda.current_stream().cuda_stream
    ref = api.Solver(mdg, pops, stream=stream)
    assert ref.kernel_path() == "quad_line"
    ref.upload(Q)
    rt_ref = ref.lsrk45(dt, nsteps, want_rhstest=True)
    Qref = ref.download()
    mds = [SD.uniform_quad_mesh_md(prd, *mesh, domain=dom, yslab=(r, nslabs)) for r in range(nslabs)]
    sol = [api.Solver(md, pops, n_halo=md.n_halo, stream=stream) for md in mds]
    e0 = 0
    for s, md in zip(sol, mds):
        assert s.kernel_path() == "quad_line"
        s.upload([np.asfortranarray(q[:, e0:e0 + md.K]) for q in Q])
        e0 += md.K
    st = H.LocalSlabs(sol, mds, torch, torch.device("cuda", 0), overlap=overlap)
    for k in range(nsteps):
        st.step(dt, want_rhstest=(k == nsteps - 1))
    Qs = [s.download() for s in sol]
    full = [np.concatenate([q[f] for q in Qs], axis=1) for f in range(4)]
    rt = sum(s.read_rhstest() for s in sol)
    diff = max(np.abs(a - b).max() for a, b in zip(full, Qref))
    print(f"quad N={N} {mesh} {nslabs} slabs overlap={overlap}: max |dQ| {diff:.2e}, rhstest {rt:.6e} vs {rt_ref:.6e}")
    assert diff == 0.0
    assert abs(rt - rt_ref) < 1e-12 * max(1.0, abs(rt_ref))


def test_quad_line_walls_as_self_mapped_faces_and_resident_rhs(monkeypatch):
    """A non-periodic quad mesh (boundary points map to themselves, as connect_mesh leaves them) on the line kernel against the
    table-driven kernels; `rhs(None)` -- the rhs of the RESIDENT state -- equals `rhs(Q)` after upload, also for a Navier-Stokes
    context whose state lives in the Gauss-nodal basis."""
    from types import SimpleNamespace
    from esdg_cns_b200 import api, setup_dg as SD
    prd, pops = SD.reference_quad(3)
    pmd = SD.uniform_quad_mesh_md(prd, 7, 6, domain=((0, 20), (-5, 5)), periodic=(False, True))
    rho, u, v, p = SD.vortex(pmd.xq, pmd.yq)
    Q = SD.primitive_to_conservative(rho, (u, v), p)
    s = api.Solver(pmd, pops)
    assert s.kernel_path() == "quad_line"
    r1, rt1 = s.rhs(Q, True)
    s.upload(Q)
    r2, rt2 = s.rhs(None, True)
    assert all(np.array_equal(a, b) for a, b in zip(r1, r2)) and rt1 == rt2
    monkeypatch.setenv("ESDG_FORCE_GENERIC", "1")
    g = api.Solver(pmd, pops)
    assert g.kernel_path() == "generic"
    r3, rt3 = g.rhs(Q, True)
    monkeypatch.delenv("ESDG_FORCE_GENERIC")
    sc = max(np.abs(x).max() for x in r3)
    d = max(np.abs(a - b).max() for a, b in zip(r1, r3)) / sc
    print(f"non-periodic quads: line vs generic {d:.2e}, rhstest {rt1:.6e} / {rt3:.6e}")
    assert d < 5e-12 and abs(rt1 - rt3) < 1e-11 * max(1.0, abs(rt3))
    # Navier-Stokes context: resident rhs in the nodal basis == rhs of the uploaded coefficients
    rdm, opm = SD.reference_quad_modal(3)
    mdm = SD.uniform_quad_mesh_md(rdm, 5, 5, periodic=(False, False))
    bc = SD.cavity_bc_sets(mdm)
    st = SimpleNamespace(Re=100.0, mu=.01, lam=-2 / 300, Pr=.71, bctype=1, inviscid_dissp=True, viscous_dissp=True,
                         wall=bc.wall, lid=bc.lid, vlid=bc.vlid, mapB=bc.mapB)
    x, y = mdm.x, mdm.y
    Qm = SD.primitive_to_conservative(1 + .1 * np.sin(np.pi * x) * np.cos(np.pi * y), (.2 * np.sin(np.pi * x) * np.sin(np.pi * y), .1 * (1 - y * y) * np.cos(x)),
                                      (1 / (.09 * 1.4)) * (1 + .05 * np.cos(np.pi * x)))
    c = api.Solver(mdm, opm, cns=st)
    assert c.kernel_path() == "cns_quad_line"
    a1 = c.rhs(Qm, True)
    c.upload(Qm)
    a2 = c.rhs(None, True)
    back = c.download()
    scm = max(np.abs(x_).max() for x_ in a1[0])
    assert max(np.abs(p_ - q_).max() for p_, q_ in zip(a1[0], a2[0])) / scm < 1e-13
    assert abs(a1[1] - a2[1]) < 1e-12 and abs(a1[2] - a2[2]) < 1e-12
    assert max(np.abs(p_ - q_).max() for p_, q_ in zip(back, Qm)) < 1e-13      # Vq then Pq: the caller's coefficients come back
