"""tcgen05 building blocks and the fused bf16 MLP kernel (sm_100a only)."""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from _cases import T

pytestmark = pytest.mark.gpu
DEV = "cuda"


def bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("N,K", [(256, 64), (256, 256), (128, 256), (64, 128), (32, 64)])
def test_umma_selftest(cg, N, K):
    """descriptor / swizzle / TMEM-layout check: C = A B^T on the tensor cores
    from thread-written shared-memory operands equals the fp32 product of the
    bf16-rounded inputs."""
    from nerf_projects_b200 import ops, _lib
    gen = torch.Generator().manual_seed(N * 7 + K + cg)
    A = torch.randn(128 * cg, K, generator=gen)
    B = torch.randn(N, K, generator=gen)
    try:
        C = ops.umma_selftest(A.to(DEV), B.to(DEV), cta_group=cg)
        torch.cuda.synchronize()
    except Exception:
        print("watchdog:", _lib.watchdog_record())
        raise
    ref = bf16r(A).double() @ bf16r(B).double().T
    err = (C.cpu().double() - ref).abs().max().item()
    assert err < 1e-3, f"max err {err}"
