"""GPU: the tcgen05 / TMEM GEMM building block (3xTF32 split operands) against an fp64 reference."""
import pytest
import torch

from hamiltorch_b200 import engine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('shape', [(128, 128, 32), (256, 384, 1024), (128, 1024, 96)])
def test_gemm_nt_tf32x3_is_fp32_accurate(shape):
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).cuda()
    B = torch.randn(N, K, generator=g).cuda()
    D = engine.gemm_nt(A, B)
    torch.cuda.synchronize()
    ref = (A.double() @ B.double().t())
    err = (D.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    fp32 = ((A @ B.t()).double() - ref).abs().max().item()       # what a plain fp32 GEMM achieves
    # 3xTF32 (hi*hi + hi*lo + lo*hi, the lo*lo term dropped): 22 significant bits per operand -> a few times the
    # rounding error of a plain fp32 GEMM, three orders of magnitude below a single tf32 product (~1e-3 relative)
    assert err <= 2e-5 * scale, (err, fp32, scale)
