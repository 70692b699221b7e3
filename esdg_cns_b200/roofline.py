"""Roofline bookkeeping of the hot path: ALGORITHMIC bytes and flops per element-stage.

These are the contract numbers `bench.py` divides by the measured kernel time; they describe the reference
algorithm, not this implementation:

  B_alg = 8 Nf (5 Ns + 2 Nfq)           Q and res read + written, traces written + read (Euler)
        + 8 Nf (Ns + (1 + 2 d) Nfq)     CNS: one more Q read, one more v-trace read, sigma traces written + read
  F_alg = P_sym F_pair + Nfq F_surf + 2 Nf (nnz Vq + nnz Ef|VhP + nnz Ph + nnz L)
        + Nq (3d+10) + Nproj (3d+14) + Nh (3d+5) + 6 Ns Nf            (FMA = 2 flops; unordered pairs once)

Audit (round 2, frozen): the survey's hand count F_pair = 106 / F_surf = 114 (3-D) was checked against what the
tensor-product kernels EXECUTE on the bench workload (ncu --set full, profiles/r2_hex_stage_ws_ncu_summary.txt:
2334 M warp instructions per launch over 262 144 elements, 22.7 % DMUL + 9.2 % DADD + 19.9 % DFMA):
(0.227 + 0.092 + 2 x 0.199) x 2334e6 x 32 / 262144 = 204 K executed flops per element-stage against
F_alg = 213 275 -- within 4 %, so the constants stay as the survey wrote them.
"""
import hashlib
import json
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)


def sizes(elem, N):
    n = N + 1
    if elem == "hex":
        return dict(d=3, nf=5, Ns=n ** 3, Nq=n ** 3, Nfq=6 * n * n, collocated=True)
    if elem == "quad":
        return dict(d=2, nf=4, Ns=n * n, Nq=n * n, Nfq=4 * n, collocated=True)
    if elem == "quad_modal":      # CNS quads: modal basis, (N+2)^2-point Gauss rule, dense operators
        nq1 = N + 2
        return dict(d=2, nf=4, Ns=n * n, Nq=nq1 * nq1, Nfq=4 * nq1, collocated=False)
    if elem == "tri":
        Np = n * (n + 1) // 2
        nq = {1: 3, 2: 6, 3: 12, 4: 16}.get(N)      # degree-2N cubature of the reference (QuadratureData/*.txt)
        return dict(d=2, nf=4, Ns=Np, Nq=nq, Nfq=3 * n, collocated=False)
    raise ValueError(elem)


def alg_bytes_sizes(nf, Ns, Nfq, d, cns=False):
    b = 8 * nf * (5 * Ns + 2 * Nfq)
    if cns:
        b += 8 * nf * (Ns + (1 + 2 * d) * Nfq)
    return b


def alg_bytes_curved_hex(N):
    """Curved hexes: the affine figure + the nine geometric factors at every hybridized point (9 Nh 8 B)."""
    s = sizes("hex", N)
    return alg_bytes_sizes(s["nf"], s["Ns"], s["Nfq"], 3) + 9 * (s["Nq"] + s["Nfq"]) * 8


def alg_bytes(elem, N, cns=False):
    s = sizes(elem, N)
    return alg_bytes_sizes(s["nf"], s["Ns"], s["Nfq"], s["d"], cns)


def alg_flops(elem, N):
    """Euler only (the CNS flop count is not part of the contract)."""
    s = sizes(elem, N)
    d, nf, Ns, Nq, Nfq = s["d"], s["nf"], s["Ns"], s["Nq"], s["Nfq"]
    n = N + 1
    Nh = Nq + Nfq
    F_pair, F_surf = (106, 114) if d == 3 else (70, 85)
    if elem == "hex":
        P_sym = 3 * n * n * (n * (n - 1) // 2) + Nfq * n
        nnz = Nfq * n + Ns * (1 + 2 * d) + Ns * 2 * d
        Nproj = Nfq
    elif elem == "quad":
        P_sym = 2 * n * (n * (n - 1) // 2) + Nfq * n
        nnz = Nfq * n + Ns * (1 + 2 * d) + Ns * 2 * d
        Nproj = Nfq
    else:                          # dense hybridized operators: Nh (Nh - 1) / 2 minus the face-face block
        P_sym = Nh * (Nh - 1) // 2 - Nfq * (Nfq - 1) // 2
        nnz = Nq * Ns + Nh * Nq + Ns * Nh + Ns * Nfq
        Nproj = Nh
    return (P_sym * F_pair + Nfq * F_surf + 2 * nf * nnz + Nq * (3 * d + 10) + Nproj * (3 * d + 14)
            + Nh * (3 * d + 5) + 6 * Ns * nf)


def kernel_source_hash():
    """sha1 over the CUDA sources: stamps ncu-derived numbers (profiles/traffic.json) with the code they were taken from."""
    h = hashlib.sha1()
    src = os.path.join(_HERE, "csrc")
    for name in sorted(os.listdir(src)):
        if name.endswith((".cu", ".cuh", ".h", ".inc")):
            with open(os.path.join(src, name), "rb") as fh:
                h.update(name.encode())
                h.update(fh.read())
    return h.hexdigest()[:16]


def measured_traffic(kernel, elements):
    """DRAM bytes per launch of `kernel` from the committed ncu capture -- only if it was taken from the current
    kernel sources and at this element count; otherwise None (a stale capture is not reported)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, "no capture"
    with open(p) as fh:
        t = json.load(fh).get(kernel)
    if not t or t.get("elements") != elements:
        return None, "no capture for this kernel / size"
    if t.get("kernel_source_hash") != kernel_source_hash():
        return None, f"stale: captured from sources {t.get('kernel_source_hash')}, now {kernel_source_hash()}"
    return t["dram_bytes_per_launch"], t.get("source", "ncu --set full")
