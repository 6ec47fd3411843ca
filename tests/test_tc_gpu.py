"""tcgen05 / TMEM / TMA building blocks on hardware: one-tile GEMM vs torch with bf16-rounded operands."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def pack_tiled_bf16(W):
    """[R, K] fp32 -> bf16 bytes in the core-matrix tiled layout of csrc/tc.cuh (chunk-major, 16 B per row)."""
    R, K = W.shape
    return W.to(torch.bfloat16).reshape(R, K // 8, 8).permute(1, 0, 2).contiguous()


@pytest.mark.parametrize("N,K", [(32, 32), (128, 32), (32, 128), (96, 64), (256, 16)])
@pytest.mark.parametrize("bulk", [False, True])
def test_umma_tile_gemm(N, K, bulk):
    from aline_b200 import _lib
    torch.manual_seed(N * 1000 + K)
    A = torch.randn(128, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    D = torch.zeros(128, N, device="cuda")
    Bp = pack_tiled_bf16(B) if bulk else None
    _lib.check(_lib.lib().aline_tc_selftest(_lib.dptr(A), _lib.dptr(B), N, K, _lib.dptr(D),
                                            ctypes.c_void_p(Bp.data_ptr()) if bulk else None, _lib.stream_ptr("cuda")))
    torch.cuda.synchronize()
    ref = A.to(torch.bfloat16).float() @ B.to(torch.bfloat16).float().T
    assert torch.allclose(D, ref, rtol=1e-4, atol=1e-3), (D - ref).abs().max().item()


@pytest.mark.parametrize("N,K", [(32, 128), (64, 32), (96, 48)])
def test_umma_tile_gemm_a_in_tmem(N, K):
    """A operand read from tensor memory (the fast query stream keeps its softmax / MLP activations there)."""
    from aline_b200 import _lib
    torch.manual_seed(N * 1000 + K + 1)
    A = torch.randn(128, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    D = torch.zeros(128, N, device="cuda")
    _lib.check(_lib.lib().aline_tc_selftest_tmem_a(_lib.dptr(A), _lib.dptr(B), N, K, _lib.dptr(D), _lib.stream_ptr("cuda")))
    torch.cuda.synchronize()
    ref = A.to(torch.bfloat16).float() @ B.to(torch.bfloat16).float().T
    assert torch.allclose(D, ref, rtol=1e-4, atol=1e-3), (D - ref).abs().max().item()
