"""Multi-GPU path: z-slab partition + NCCL halo exchange must reproduce the single-GPU result."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("overlap,generic,N", [(1, 0, 2), (0, 0, 3), (1, 0, 4), (1, 1, 2)])
def test_two_gpu_slabs_match_single_gpu(overlap, generic, N):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(29530 + 3 * N + overlap + 2 * generic),
           os.path.join(ROOT, "tests", "dist_check.py"), "--overlap", str(overlap), "--generic", str(generic),
           "--N", str(N), "--K", "4", "3", "6"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0


@pytest.mark.gpu
@pytest.mark.parametrize("overlap,generic,N,cns", [(1, 0, 4, 0), (0, 0, 3, 0), (1, 1, 2, 0), (1, 0, 3, 1)])
def test_two_gpu_in_library_nccl_matches_single_gpu(overlap, generic, N, cns):
    """esdg_comm_init: the context owns its NCCL communicator; esdg_lsrk45_step exchanges the halos itself (ncclSend from
    one contiguous message per neighbour and plane, ncclRecv straight into the halo tail, rhstest all-reduced)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + 5 * N + overlap + 2 * generic + 3 * cns),
           os.path.join(ROOT, "tests", "dist_check.py"), "--ccomm", "1", "--overlap", str(overlap), "--generic", str(generic),
           "--N", str(N), "--cns", str(cns), "--K"] + (["6", "4", "0"] if cns else ["4", "3", "6"])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0


@pytest.mark.gpu
def test_two_gpu_cns_slabs_match_single_gpu():
    """Distributed CNS (two NCCL halo exchanges per stage) on 2 GPUs against the single-GPU run."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29571",
           os.path.join(ROOT, "tests", "dist_check.py"), "--cns", "1", "--N", "3", "--K", "6", "4", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0


@pytest.mark.gpu
@pytest.mark.parametrize("ccomm,overlap", [(1, 1), (1, 0), (0, 1)])
def test_two_gpu_quad_slabs_match_single_gpu(ccomm, overlap):
    """Euler on y-slabs of a quad mesh (tensor-product quad kernels, 7-plane halo): in-library NCCL exchange overlapped with
    the interior rows, without overlap, and host-driven -- bit-identical to the single-GPU run."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(29650 + 2 * ccomm + overlap),
           os.path.join(ROOT, "tests", "dist_check.py"), "--quad", "1", "--ccomm", str(ccomm), "--overlap", str(overlap),
           "--N", "3", "--K", "6", "8", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0
