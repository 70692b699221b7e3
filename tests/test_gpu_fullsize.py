"""Size-independent properties at BASELINE.json's full sizes (the oracle cannot run these in seconds):
config 4 (hex N=3, 64^3 elements) and the per-GPU slab of config 5 (hex N=4, 128 x 128 x 16 elements).

* free-stream preservation: rhs of a constant state vanishes (metric identities + consistency of the flux);
* discrete entropy conservation: with the Lax-Friedrichs penalty off, sum_f sum(wJq .* v_f .* rhs_f) is zero to
  round-off on the reference's own random state (dg3D_euler_hex.jl:105-110, :224-226); with it on it is negative;
* conservation: sum(wJq .* rhs_f) = 0 for every field on the periodic mesh;
* the fused LSRK stage is bitwise reproducible run to run.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,mesh", [(3, (64, 64, 64)), (4, (128, 128, 16))])
def test_full_size_properties(N, mesh):
    from esdg_cns_b200 import api, setup_dg as SD
    prd, pops = SD.reference_hex(N, lf_scale=0.0)
    pmd = SD.uniform_hex_mesh_md(prd, mesh, domain=((-np.pi, np.pi),) * 3)
    n = N + 1
    K = pmd.K
    w1 = np.asarray(prd.wq).reshape(-1)
    wJ = w1 * float(np.asarray(pmd.J).reshape(-1)[0])          # uniform mesh: one J
    s = api.Solver(pmd, pops)
    # free stream
    one = np.ones((n ** 3, K))
    Qc = SD.primitive_to_conservative(1.2 * one, (.3 * one, -.2 * one, .1 * one), .9 * one)
    out, _ = s.rhs(Qc, False)
    fs = max(float(np.abs(x).max()) for x in out)
    del out, Qc
    # entropy conservation on the reference's random state
    rng = np.random.default_rng(0)
    rho = 2 + .1 * rng.random((n ** 3, K))
    p = 1 + .1 * rng.random((n ** 3, K))
    Q = SD.primitive_to_conservative(rho, (0 * one, one, 0 * one), p)
    del rho, p
    out, rt = s.rhs(Q, True)
    scale = sum(float(np.abs(x).max()) for x in out)
    cons = [float(wJ @ x.sum(axis=1)) for x in out]
    vol = float(wJ.sum()) * K
    print(f"N={N} mesh={mesh}: free stream {fs:.2e}; EC rhstest {rt:.3e} (|rhs| ~ {scale:.2e}, volume {vol:.1f}); "
          f"conservation {max(abs(c) for c in cons):.2e}")
    assert fs < 1e-11
    assert abs(rt) < 1e-12 * scale * vol                      # round-off of a sum over K*Nq*5 terms
    assert max(abs(c) for c in cons) < 1e-12 * scale * vol
    del out
    s.close()
    # dissipation with the penalty on, and run-to-run reproducibility of two LSRK steps
    prd, pops = SD.reference_hex(N, lf_scale=.25)
    s = api.Solver(pmd, pops)
    res = []
    for _ in range(2):
        s.upload(Q)
        rt_lf = s.lsrk45(1e-4, 2, want_rhstest=True)
        res.append((s.download(), rt_lf))
    assert res[0][1] < 0 and res[0][1] == res[1][1]
    assert all(np.array_equal(a, b) for a, b in zip(res[0][0], res[1][0]))
    assert all(np.isfinite(a).all() for a in res[0][0])
    s.close()
