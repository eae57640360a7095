"""One GPU with forced ghost slabs (DDA_FORCE_SPLIT: the rank is its own periodic neighbour, so the partitioned kernel
variants -- D_W interior / boundary modes, interior / boundary SAP block lists, ghost branch of the coarse combine, pack
kernel, second-stream exchange -- run and are compared with the oracle on a single-GPU box), and two / four GPUs, lattice split along T, NCCL halos + allreduce through the CUDA library: every rank checks its part
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


@pytest.mark.parametrize("dirs,levels,lattice", [("T", 2, None), ("TZ", 3, None), ("T", 3, "16,16,16,16"), ("TZ", 2, "16,16,8,8")])
def test_one_gpu_forced_split(cuda_lib, oracle_ref, tmp_path, dirs, levels, lattice):
    """lattice given: synthetic field, 4^4 blocks, 16 sites in the split directions -> a rank owns interior AND boundary
    blocks, so the fused fine SAP kernel runs its interior list during the exchange and the boundary list after it."""
    check(run_ranks(cuda_lib, "nccl", levels, tmp_path, world=1, lattice=lattice, extra_env={"DDA_FORCE_SPLIT": dirs}))
