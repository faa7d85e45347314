"""Multi-GPU parity (needs >= 2 B200s; skipped otherwise): partitioned assembly + NCCL halo sums + Krylov solve
against the single-partition oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    # not torch.cuda.device_count(): importing torch into a process that already holds the reference's
    # libsvref.so with RTLD_GLOBAL (tests/test_gpu_hostshim.py) resolves torch symbols into it and crashes
    try:
        out = subprocess.run(["nvidia-smi", "-L"], stdout=subprocess.PIPE, text=True, timeout=60).stdout
    except Exception:
        return 0
    return sum(1 for line in out.splitlines() if line.startswith("GPU "))


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
@pytest.mark.parametrize("mode,ls", [("slab", "gmres"), ("scattered", "gmres"), ("slab", "ns"), ("fsi", "gmres+cg"), ("metis", "gmres")])
def test_two_gpu_parity(mode, ls, transport):
    """transport p2p: shared-node sums and scalar all-reduces by the library's own kernels over peer memory (CUDA IPC
    mailboxes, NVLink); nccl: ncclSend/Recv + ncclAllReduce.  Same parity bar for both."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    if mode == "metis" and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libsvmetis.so")):
        pytest.skip("needs oracle/_ref/libsvmetis.so (make -C oracle metis)")
    n = min(_ngpu(), 4) if mode in ("scattered", "fsi") else (_ngpu() if mode == "metis" else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "mgpu_worker.py"), mode, ls]
    env = dict(os.environ)
    env["SVB200_COMM"] = transport
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, env=env)
    print(r.stdout[-2000:])
    assert r.returncode == 0, r.stdout[-4000:]
    assert f"transport {transport}" in r.stdout, "the requested transport was not the one used:\n" + r.stdout[-2000:]


@pytest.mark.parametrize("mode", ["slab", "metis"])
def test_multi_gpu_rank_local_parity(mode):
    """Every GPU rank against the SAME rank of a multi-rank run of the compiled reference (N processes of libsvref.so over the
    shared-memory MPI shim): identical local CSR, lhs.map / shared-node lists, local Val and R (after the shared-node sum) to
    1e-12, the GMRES solution to 1e-6 — no gluing and no single-partition stand-in."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libsvref.so")):
        pytest.skip("needs oracle/_ref/libsvref.so")
    if mode == "metis" and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libsvmetis.so")):
        pytest.skip("needs oracle/_ref/libsvmetis.so")
    n = 2 if mode == "slab" else min(_ngpu(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29543", os.path.join(ROOT, "tests", "mgpu_worker.py"), mode, "gmres", "ranklocal"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, env=dict(os.environ))
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-4000:]
