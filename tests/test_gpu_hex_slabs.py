"""Single-GPU coverage of the distributed Euler-hex path: the mesh split into z-slabs, one context
per slab with halo slots, driven in one process by halo.LocalSlabs -- the same esdg_stage_update parts
(boundary layers first, then the interior), pack / unpack and 7+-plane halo the NCCL stepper uses; only the
transport is a device-to-device copy.  Must reproduce the single-context run bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,mesh,nslabs,overlap,generic", [
    (4, (5, 4, 6), 2, True, False), (4, (4, 3, 9), 3, True, False), (3, (6, 5, 6), 3, True, False),
    (3, (5, 4, 4), 2, False, False), (4, (4, 4, 3), 3, True, False),      # one layer per slab: no interior part
    (2, (4, 3, 6), 2, True, True), (1, (7, 5, 6), 3, True, False)])
def test_hex_slabs_with_halo_exchange_match_single_context(N, mesh, nslabs, overlap, generic, monkeypatch):
    import torch
    from esdg_cns_b200 import api, halo as H, setup_dg as SD
    if generic:
        monkeypatch.setenv("ESDG_FORCE_GENERIC", "1")
    prd, pops = SD.reference_hex(N)
    dom = ((-np.pi, np.pi),) * 3
    mdg = SD.uniform_hex_mesh_md(prd, mesh, domain=dom)
    x, y, z = SD.hex_node_coords(prd, mdg)
    rho, u, v, w, p = SD.taylor_green(x, y, z)
    Q = SD.primitive_to_conservative(rho + .05 * np.sin(x) * np.cos(z), (u, v, w + .1 * np.sin(y + z)), p)
    dt, nsteps = 1e-3, 2
    stream = torch.cuda.current_stream().cuda_stream
    ref = api.Solver(mdg, pops, stream=stream)
    ref.upload(Q)
    rt_ref = ref.lsrk45(dt, nsteps, want_rhstest=True)
    Qref = ref.download()
    mds = [SD.uniform_hex_mesh_md(prd, mesh, domain=dom, zslab=(r, nslabs)) for r in range(nslabs)]
    sol = [api.Solver(md, pops, n_halo=md.n_halo, stream=stream) for md in mds]
    e0 = 0
    for s, md in zip(sol, mds):
        s.upload([np.asfortranarray(q[:, e0:e0 + md.K]) for q in Q])
        e0 += md.K
    st = H.LocalSlabs(sol, mds, torch, torch.device("cuda", 0), overlap=overlap)
    for k in range(nsteps):
        st.step(dt, want_rhstest=(k == nsteps - 1))
    Qs = [s.download() for s in sol]
    full = [np.concatenate([q[f] for q in Qs], axis=1) for f in range(5)]
    rt = sum(s.read_rhstest() for s in sol)
    diff = max(np.abs(a - b).max() for a, b in zip(full, Qref))
    print(f"hex N={N} {mesh} {nslabs} slabs overlap={overlap} generic={generic}: max |dQ| {diff:.2e}, rhstest {rt:.6e} vs {rt_ref:.6e}")
    assert diff == 0.0
    assert abs(rt - rt_ref) < 1e-12 * max(1.0, abs(rt_ref))


@pytest.mark.parametrize("N,mesh,periodic,nslabs", [(4, (5, 4, 6), (True, True, True), 1), (3, (4, 3, 5), (True, False, True), 1),
                                                    (2, (6, 5, 4), (False, False, False), 1), (4, (4, 3, 6), (True, True, True), 3),
                                                    (2, (4, 4, 4), (True, True, False), 2)])
def test_structured_hex_setup_on_the_device_matches_host_built_maps(N, mesh, periodic, nslabs):
    """esdg_set_structured_hex: mapP and the geometric factors generated on the device give bitwise
    the states of the host-built arrays of uniform_hex_mesh_md -- whole meshes (periodic and with physical boundaries)
    and z-slabs with halo slots."""
    import torch
    from esdg_cns_b200 import api, halo as H, setup_dg as SD
    prd, pops = SD.reference_hex(N)
    dom = ((-np.pi, np.pi), (-2.0, 1.0), (0.0, 3.0))
    res = []
    for structured in (False, True):
        make = SD.structured_hex_md if structured else SD.uniform_hex_mesh_md
        stream = torch.cuda.current_stream().cuda_stream
        if nslabs == 1:
            md = make(prd, mesh, domain=dom, periodic=periodic)
            x, y, z = SD.hex_node_coords(prd, md)
            rho, u, v, w, p = SD.taylor_green(x, y, z)
            Q = SD.primitive_to_conservative(rho + .05 * np.sin(x), (u + .3, v - .2, w + .1 * np.sin(y) + .4), p)
            s = api.Solver(md, pops, stream=stream)
            assert s.hex_diag()
            s.upload(Q)
            rt = s.lsrk45(1e-3, 2, want_rhstest=True)
            res.append((np.stack(s.download()), rt))
        else:
            mds = [make(prd, mesh, domain=dom, periodic=periodic, zslab=(r, nslabs)) for r in range(nslabs)]
            sol = [api.Solver(md, pops, n_halo=md.n_halo, stream=stream) for md in mds]
            for s, md in zip(sol, mds):
                x, y, z = SD.hex_node_coords(prd, md)
                rho, u, v, w, p = SD.taylor_green(x, y, z)
                s.upload(SD.primitive_to_conservative(rho + .05 * np.sin(x), (u + .3, v - .2, w + .1 * np.sin(y) + .4), p))
            st = H.LocalSlabs(sol, mds, torch, torch.device("cuda", 0))
            for k in range(2):
                st.step(1e-3, want_rhstest=(k == 1))
            res.append((np.concatenate([np.stack(s.download()) for s in sol], axis=2), sum(s.read_rhstest() for s in sol)))
    assert np.array_equal(res[0][0], res[1][0])
    assert res[0][1] == res[1][1]
