"""GPU parity in the regime the benchmark runs in.

bench.py runs Taylor-Green at Ma = 0.1 on 128^3 hexes, N = 4: node spacing h ~ 2 pi / 128 / 5, where almost every
pair of the flux-differencing sum takes the TAYLOR branch of the reference's logmean (logmean.jl:17-25) and a few
per cent sit next to its |f| < 1e-4 switch.  The coarse-mesh parity tests (test_gpu_parity.py) exercise the log
branch almost exclusively, so here the tensor-product kernels are compared on

  * the bench state itself: a K1D^3 sub-box of the 128^3 mesh (same h, same per-node variation), N = 4 and 3,
  * a well-conditioned all-Taylor state (rho, p vary by 1e-5, velocities O(1)): the north star's 1e-12 must hold,
  * a state whose pairs straddle the switch,

against BOTH float64 oracles (numpy: oracle/solvers.py, scalar C: oracle/c) and against the extended-precision
evaluation of the same formulas (oracle/extended.py).  At Ma = 0.1 on this mesh the rhs is a small difference of
O(p) fluxes: any two float64 evaluations of the reference formulas differ by ~1e-11 relative through rounding
alone (the two oracles do), so the bound there is "no further from the 80-bit value than the float64 oracles are".
"""
import numpy as np
import pytest

from helpers import relerr as _relerr, NORTH_STAR_TOL, noise_floor

pytestmark = pytest.mark.gpu

H_BENCH = 2 * np.pi / 128          # element size of BASELINE config 5


def _subbox(N, K1D, shift=(0.7, -0.4, 0.3)):
    """K1D^3 periodic box of bench-sized elements, centred at `shift` of the Taylor-Green domain."""
    from oracle import solvers as SV
    s = .5 * K1D * H_BENCH
    rd, md, ops = SV.setup_hex(N, K1D, scale=(s, s, s), shift=shift)
    return rd, _exactly_affine(md, rd), ops


def _exactly_affine(md, rd):
    """The oracle's setup differentiates the node coordinates (geometric_factors.jl:34-67), which on a box of 0.05-wide
    elements away from the origin leaves 2.5e-11 relative noise in rxJ.. from node to node.  The reference's hex
    script averages the per-node values pair by pair (dg3D_euler_hex.jl:145-151), so that noise breaks the metric
    identities and moves ITS rhs by 3.5e-10 here (printed by the test below) -- input noise, not arithmetic.  The
    tensor-product kernels and the C oracle take one value per affine element (row [1,e], as the 2-D scripts do), so
    for a like-for-like comparison every consumer gets the exactly-affine arrays."""
    for n in ("rxJ", "sxJ", "txJ", "ryJ", "syJ", "tyJ", "rzJ", "szJ", "tzJ", "J"):
        a = getattr(md, n)
        a[:] = 0.0 if n in ("sxJ", "txJ", "ryJ", "tyJ", "rzJ", "szJ") else a[:1]     # axis-aligned box: exact zeros
    md.wJq = np.diag(rd.wq) @ md.J
    nfp = md.nxJ.shape[0] // 6
    for n in ("nxJ", "nyJ", "nzJ", "sJ"):
        a = getattr(md, n)
        for f in range(6):
            a[f * nfp:(f + 1) * nfp] = a[f * nfp:f * nfp + 1]
        if n != "sJ":
            # 3e-11 sJ of noise in the tangential normal components would meet the reference's sqrt(|u_n|)
            # (euler_variables.jl:9) wherever the flow is aligned with the mesh (w = 0 in Taylor-Green)
            a[np.abs(a) < 1e-8 * md.sJ] = 0.0
    return md


def _states(md, which):
    from oracle import physics as P
    x, y, z = md.xq, md.yq, md.zq
    if which in ("tgv", "tgv_w"):                    # the bench IC on its own mesh (bench.py)
        rho, u, v, w, p = P.taylor_green(x, y, z)
        if which == "tgv_w":                         # ... with a vertical velocity, so that no u_n is at round-off level
            w = .3 * np.sin(x + y + z)
    else:
        k = 2 * np.pi / (md.xq.max() - md.xq.min() + 1e-300) * (1 - 1e-3)
        sx, cy, cz = np.sin(k * x), np.cos(k * y + .3), np.cos(k * z - .2)
        amp = 1e-5 if which == "taylor" else 1.2e-3  # neighbour-to-neighbour f ~ 1e-6 / ~1e-4
        rho = 1 + amp * sx * cy * cz
        p = 1 + amp * np.cos(k * x + .5) * np.sin(k * y) * cz
        u = .6 + .3 * sx * cy
        v = -.4 + .3 * np.cos(k * x) * np.sin(k * z)
        w = .5 + .2 * np.sin(k * (x + y + z))
    return list(P.primitive_to_conservative(rho, (u, v, w), p))


def _pair_stats(Q, md, ops, rd):
    """Fraction of (rho, beta) pairs of the Hadamard sum in the Taylor branch / within 10 % of the switch."""
    from oracle import physics as P
    VU = P.v_of_u(Q)
    Uf = P.u_of_v(tuple(ops.Ef @ v for v in VU))
    Uh = tuple(np.vstack([q, uf]) for q, uf in zip(Q, Uf))
    out = {}
    for name, a in (("rho", Uh[0]), ("beta", P.beta_of_u(Uh))):
        fs = []
        for i, js in enumerate(ops.Qnzids):
            for j in js:
                if j > i:
                    fs.append(np.abs(2 * (a[j] - a[i]) / (a[j] + a[i])))
        f = np.concatenate(fs)
        out[name] = (float((f < 1e-4).mean()), float(((f > .9e-4) & (f < 1.1e-4)).mean()))
    return out


@pytest.mark.parametrize("N,K1D", [(4, 6), (3, 6)])
@pytest.mark.parametrize("which", ["tgv", "taylor", "straddle"])
@pytest.mark.parametrize("lf", [0.0, 0.25])
def test_hex_rhs_at_bench_resolution(N, K1D, which, lf):
    from oracle import solvers as SV, extended as X
    from oracle.c_oracle import COracle
    from esdg_cns_b200 import api
    rd, md, ops = _subbox(N, K1D)
    ops.lf_scale = lf
    Q = _states(md, which)
    st = _pair_stats(Q, md, ops, rd)
    ref, rt_ref = SV.rhs_collocated(Q, md, ops, rd, lf, True)
    cref, _ = COracle(md, ops, lf_scale=lf).rhs(Q, False)
    truth, _ = X.rhs_collocated_ld(Q, md, ops, rd, lf, False)
    truth = [np.asarray(t, dtype=np.float64) for t in truth]
    s = api.Solver(md, ops)
    assert s.kernel_time_ms("generic_rhs")[1] == 0 and s.trace_planes() >= 7      # tensor-product kernels
    assert s.hex_diag()                                                           # ... in the axis-aligned form the bench runs
    out, rt = s.rhs(Q, True)
    e_gpu_np, e_gpu_c = _relerr(out, ref), _relerr(out, cref)
    e_gpu_t, e_np_t, e_c_t = _relerr(out, truth), _relerr(ref, truth), _relerr(cref, truth)
    cloud = max(e_np_t, e_c_t, _relerr(ref, cref))
    print(f"bench-regime hex N={N} {K1D}^3 {which} lf={lf}: Taylor fraction rho {st['rho'][0]:.3f} beta {st['beta'][0]:.3f}, "
          f"within 10% of the switch rho {st['rho'][1]:.3f} beta {st['beta'][1]:.3f}; "
          f"gpu-numpy {e_gpu_np:.2e} gpu-C {e_gpu_c:.2e} numpy-C {_relerr(ref, cref):.2e}; "
          f"vs 80-bit: gpu {e_gpu_t:.2e} numpy {e_np_t:.2e} C {e_c_t:.2e}")
    if which == "tgv":       # whole 128^3 field: 100 % / 90 %; this sub-box sits where the pressure gradient is large
        assert st["rho"][0] > .9 and st["beta"][0] > .5 and st["beta"][1] > .02
    elif which == "taylor" and N == 4:
        assert st["rho"][0] == 1.0 and st["beta"][0] == 1.0       # nothing but the Taylor branch
        # (N = 3: the entropy-projected face densities differ from the nodal ones by 3e-4, so a quarter of the
        #  volume-face pairs sit just above the switch -- the well-conditioned LOG branch at rho ~ 1, see hex_common.cuh)
    else:
        assert st["rho"][1] + st["beta"][1] > .02                 # pairs next to the switch exist
    # float64 rounding cloud of the reference formulas on this state: the kernel must sit inside it
    assert e_gpu_t < max(NORTH_STAR_TOL, 2.0 * cloud)
    if which == "taylor":
        assert e_gpu_np < NORTH_STAR_TOL and e_gpu_c < NORTH_STAR_TOL   # well conditioned: the north star's 1e-12
    assert abs(rt - rt_ref) < 1e-10 * max(1.0, abs(rt_ref))
    if lf == 0.0:
        assert abs(rt) < 1e-10
    else:
        assert rt <= 1e-12


@pytest.mark.parametrize("N,K1D", [(4, 6), (3, 8)])
def test_hex_lsrk_steps_at_bench_resolution(N, K1D):
    """Two LSRK45 steps of the bench state at the bench's own dt (bench.py: CFL .5) against both oracles."""
    from oracle import solvers as SV
    from oracle.c_oracle import COracle
    from esdg_cns_b200 import api
    rd, md, ops = _subbox(N, K1D)
    ops.lf_scale = .25
    Q = _states(md, "tgv_w")
    dt = .5 * 2 / (((N + 1) * (N + 2) * 3 / 2) * 128) * np.pi       # bench.py's dt on [-pi,pi]^3
    Qc, _ = COracle(md, ops, lf_scale=.25).lsrk45(Q, dt, 2)
    Qo, res = [q.copy() for q in Q], [np.zeros_like(q) for q in Q]
    for _ in range(2):
        Qo, res, rto = SV.lsrk45_step(Qo, res, dt, lambda q, t: SV.rhs_collocated(q, md, ops, rd, .25, t))
    s = api.Solver(md, ops)
    s.upload(Q)
    rtg = s.lsrk45(dt, 2, want_rhstest=True)
    Qg = s.download()
    e_np = max(np.abs(a - b).max() / np.abs(b).max() for a, b in zip(Qg, Qo))
    e_c = max(np.abs(a - b).max() / np.abs(b).max() for a, b in zip(Qg, Qc))
    e_oo = max(np.abs(a - b).max() / np.abs(b).max() for a, b in zip(Qc, Qo))
    print(f"bench-regime lsrk x2 N={N}: state gpu-numpy {e_np:.2e} gpu-C {e_c:.2e} numpy-C {e_oo:.2e}; rhstest {rtg:.6e} vs {rto:.6e}")
    assert e_np < 2e-13 + 2 * e_oo and e_c < 2e-13 + 2 * e_oo
    assert abs(rtg - rto) < 1e-9 * max(1.0, abs(rto))


def test_pair_flux_across_the_logmean_switch():
    """Pair-level sweep of the kernels' own two-point flux (esdg_debug_pair_flux -> pair_flux_v, both the
    axis-aligned and the general form) over f = 2(aR-aL)/(aR+aL) in [0, 3e-4], including f = 1e-4 +- a few ulp,
    against logmean.jl:14-28 / euler_fluxes.jl:51-89 as restated in oracle/physics.py (logs supplied by the
    caller, as in the Hadamard sum)."""
    from oracle import physics as P
    from esdg_cns_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(11)
    f = np.concatenate([np.linspace(0, 3e-4, 4001), 1e-4 * (1 + rng.uniform(-1e-3, 1e-3, 2000)), [0.0, 1e-4, 1e-4, -1e-4]])
    n = f.size

    def side(a0, fa):
        aL = np.full(n, a0)
        aR = aL * (2 + fa) / (2 - fa)
        return aL, aR

    worst = {}
    for case in ("rho", "beta", "both"):
        rL, rR = side(1.3, f if case in ("rho", "both") else 0 * f + 3e-3)
        bL, bR = side(.0071, (f[::-1] if case == "both" else f) if case in ("beta", "both") else 0 * f - 2e-3)
        # nudge the threshold points by -4..4 ulp
        for k in range(1, 10):
            rR[-k % n] = np.nextafter(rR[-k % n], np.inf if k % 2 else -np.inf)
        uL = [rng.uniform(-1, 1, n) for _ in range(3)]
        uR = [u + rng.uniform(-1e-3, 1e-3, n) for u in uL]
        UL = np.ascontiguousarray(np.stack([rL, *uL, bL, np.log(rL), np.log(bL)]))
        UR = np.ascontiguousarray(np.stack([rR, *uR, bR, np.log(rR), np.log(bR)]))
        F = np.empty((4, 5, n))
        assert lib.esdg_debug_pair_flux(0, UL.ctypes.data, UR.ctypes.data, n, 1.4, F.ctypes.data) == 0
        ref = P.euler_fluxes_3d(tuple(UL[:5]), tuple(UR[:5]), (UL[5], UL[6]), (UR[5], UR[6]))
        fr = np.abs(2 * (rR - rL) / (rR + rL))
        fb = np.abs(2 * (bR - bL) / (bR + bL))
        near = (np.abs(fr - 1e-4) < 4e-20 + 8 * np.spacing(1e-4)) | (np.abs(fb - 1e-4) < 8 * np.spacing(1e-4))
        err = np.zeros(n)
        for d in range(3):
            for k in range(5):
                err = np.maximum(err, np.abs(F[1 + d, k] - ref[d][k]) / np.maximum(np.abs(ref[d][k]), 1e-3))
        for k in range(5):
            err = np.maximum(err, np.abs(F[0, k] - ref[0][k]) / np.maximum(np.abs(ref[0][k]), 1e-3))
        worst[case] = (err[~near].max(), int((err[near] > 1e-12).sum()), int(near.sum()))
        # away from the last few ulp around the switch the kernel takes the reference's branch and agrees to round-off
        assert err[~near].max() < 2e-13, (case, err[~near].max(), np.argmax(np.where(near, 0, err)))
    print("pair flux across the logmean switch: max rel err (off-switch), # > 1e-12 within 8 ulp of the switch, # such points:", worst)


# ---- the 2-D configs at THEIR timing-scale node spacing (bench.py other_configs) -----------------------------------
def test_quad_at_bench_spacing():
    """bench.py times config 1 on 1024^2 quads over [0,20]x[-5,5] (elements 0.0195 x 0.0098): a 24x24 sub-box of that
    mesh around the vortex core, same element size, through the tensor-product quad kernels and -- same operators -- the
    table-driven ones.  Here nearly every pair takes the Taylor branch of logmean."""
    from oracle import solvers as SV, setupdg as S, mesh as M, physics as P
    from esdg_cns_b200 import api
    N, Kb = 3, 24
    hx, hy = 20.0 / 1024, 10.0 / 1024
    VX, VY, EToV = M.uniform_quad_mesh(Kb, Kb)
    VX = 5.0 + .5 * Kb * hx * VX
    VY = .3 + .5 * Kb * hy * VY
    rd = S.init_reference_quad(N)
    md = S.init_mesh((VX, VY), EToV, rd)
    S.make_periodic(md, rd, (Kb * hx, Kb * hy))
    ops = SV.build_sbp_ops(rd)
    SV.interp_geofacs(md, rd, ops)
    # differentiated node coordinates of 0.02-wide elements away from the origin carry ~1e-10 relative noise from node
    # to node (see _exactly_affine): every consumer gets the exactly-affine arrays (the 2-D scripts read row [1,e] anyway)
    for nm in ("rxJ", "sxJ", "ryJ", "syJ", "J"):
        a = getattr(md, nm)
        a[:] = 0.0 if nm in ("sxJ", "ryJ") else a[:1]
    md.wJq = np.diag(rd.wq) @ (rd.Vq @ md.J if md.J.shape[0] != rd.wq.size else md.J)
    nfp = md.nxJ.shape[0] // 4
    for nm in ("nxJ", "nyJ", "sJ"):
        a = getattr(md, nm)
        for f in range(4):
            a[f * nfp:(f + 1) * nfp] = a[f * nfp:f * nfp + 1]
        if nm != "sJ":
            a[np.abs(a) < 1e-8 * md.sJ] = 0.0
    rho, u, v, p = P.vortex(md.xq, md.yq, 0)
    Q = list(P.primitive_to_conservative(rho, (u, v), p))
    ref, rt_ref = SV.rhs_collocated(Q, md, ops, rd, .5, True)
    tol = 2 * max(NORTH_STAR_TOL, 4 * noise_floor(lambda: SV.rhs_collocated(Q, md, ops, rd, .5, False)[0], ref))
    s = api.Solver(md, ops)
    assert s.kernel_path() == "quad_line"
    out, rt = s.rhs(Q, True)
    err = _relerr(out, ref)
    print(f"quad sub-box of the 1024^2 bench mesh: relerr {err:.2e} (bound {tol:.2e}); rhstest {rt:.3e} / {rt_ref:.3e}")
    assert err < tol
    assert abs(rt - rt_ref) < 1e-11 * max(1.0, abs(rt_ref))


def test_cns_quad_at_bench_spacing():
    """bench.py times config 3 on 512^2 quads over [-1,1]^2 (h = 0.0039).  The oracle cannot hold that mesh, so the same
    neighbour-to-neighbour variation is put on 32^2 elements: a smooth O(1) cavity-like field scaled by 32/512."""
    from oracle import cns, physics as P
    from esdg_cns_b200 import api
    from types import SimpleNamespace
    rd, md, ops, bc, _ = cns.setup_cavity("quad", 3, 32)
    par = cns.cavity_params(Re=1000.0, bctype=1)
    x, y = md.x, md.y
    a = 32.0 / 512
    rho = 1 + a * .2 * np.sin(np.pi * x) * np.cos(np.pi * y)
    u = a * (1 - x * x) * (1 + y) * .5
    v = a * .3 * np.sin(np.pi * x) * (1 - y * y)
    p = (1 / (.3 ** 2 * 1.4)) * (1 + a * .1 * np.cos(np.pi * (x + y)))
    Q = list(P.primitive_to_conservative(rho, (u, v), p))
    ref, rt_ref, rtv_ref = cns.rhsRK(Q, md, ops, bc, par)
    tol = 2 * max(NORTH_STAR_TOL, 4 * noise_floor(lambda: cns.rhsRK(Q, md, ops, bc, par)[0], ref))
    st = SimpleNamespace(Re=par.Re, mu=par.mu, lam=par.lam, Pr=par.Pr, bctype=par.bctype, inviscid_dissp=True, viscous_dissp=True,
                         wall=bc.wall, lid=bc.lid, vlid=bc.vlid, mapB=md.mapB)
    s = api.Solver(md, ops, cns=st)
    assert s.kernel_path() == "cns_quad_line"
    out, rt, rtv = s.rhs(Q, True)
    err = _relerr(out, ref)
    print(f"CNS quads at bench-like spacing: relerr {err:.2e} (bound {tol:.2e}); rhstest {rt:.3e} / {rt_ref:.3e}, visc {rtv:.3e} / {rtv_ref:.3e}")
    assert err < tol
    assert abs(rt - rt_ref) < 1e-10 * max(1.0, abs(rt_ref)) and abs(rtv - rtv_ref) < 1e-10 * max(1.0, abs(rtv_ref))
