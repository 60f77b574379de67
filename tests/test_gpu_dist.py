"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): launches tests/dist_check.py under torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("backend", ["cholqr2", "tsqr"])
def test_two_rank_pipeline_matches_one_way_oracle(backend):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "dist_check.py"), "--grid", "32", "--backend", backend]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.parametrize("backend", ["cholqr2", "tsqr"])
def test_two_rank_scattered_exchange_powerlaw(backend):
    # C4-like: dense level-s closure, scattered send lists (pack + push), very ragged rows
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29547", os.path.join(ROOT, "tests", "dist_check.py"), "--matrix", "powerlaw", "--sstep", "4", "--blocks", "4",
           "--backend", backend]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.parametrize("extra", [["--halo-level", "4"], ["--halo-level", "1"], ["--own-rows-only"]])
def test_two_rank_multi_exchange_matrix_powers(extra):
    # ghost closure shallower than s: the MPK exchanges the current basis column every L steps (L = 4: two exchanges per block;
    # L = 1: the classic per-step exchange, what a rank that only holds its own rows of A gets)
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29553", os.path.join(ROOT, "tests", "dist_check.py"), "--grid", "32", "--backend", "cholqr2"] + extra
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]



@pytest.mark.parametrize("backend", ["tsqr", "cholqr2"])
def test_two_rank_full_reorthogonalisation_and_two_block_projection(backend):
    # the 'full' driver over 2 ranks (a second projection against ALL earlier vectors per block) and, inside dist_check, a
    # two-block projectAndNormalize whose second pass does not fire -- the branches a P-rank TSQR plan must keep apart
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29559", os.path.join(ROOT, "tests", "dist_check.py"), "--grid", "32", "--backend", backend, "--full-reorth"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
