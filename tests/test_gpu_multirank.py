"""Two GPUs, lattice split along T, NCCL halos + allreduce through the CUDA library: every rank checks its part
against the single-rank oracle (same worker as the gloo test).  Skipped on a single-GPU box."""
import pytest

from test_multirank_gloo import run_ranks, check

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("levels", [2, 3])
def test_two_gpus_split_T(cuda_lib, oracle_ref, tmp_path, levels):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    check(run_ranks(cuda_lib, "nccl", levels, tmp_path))


def test_two_gpus_split_Z(cuda_lib, oracle_ref, tmp_path):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    check(run_ranks(cuda_lib, "nccl", 2, tmp_path, world=2, grid="1,2"))


def test_four_gpus_split_T_and_Z(cuda_lib, oracle_ref, tmp_path):
    if _ngpu() < 4:
        pytest.skip("needs 4 GPUs")
    check(run_ranks(cuda_lib, "nccl", 2, tmp_path, world=4, grid="2,2"))
