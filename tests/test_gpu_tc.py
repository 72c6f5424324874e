"""GPU: tcgen05 kind::tf32 bring-up -- 3xTF32 GEMM against fp64, A operand from TMEM and from shared memory."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(K, N, mode, passes, seed=0):
    from ssf_slam_b200 import _native as nat
    from ssf_slam_b200 import tc
    nat.require_device()
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(128, K, generator=g)
    W = torch.randn(N, K, generator=g)
    hi, lo = tc.weight_image(W)
    Y = torch.full((128, N), float("nan"), device="cuda")
    Xd, hid, lod = X.cuda(), hi.cuda(), lo.cuda()  # keep the device tensors alive across the launch
    nat.check(nat.lib().ssf_tc_gemm_test(nat.ptr(Xd), nat.ptr(hid), nat.ptr(lod), K, N, mode, passes, nat.ptr(Y), nat.stream()))
    torch.cuda.synchronize()
    ref = (X.double() @ W.double().t())
    err = float((Y.cpu().double() - ref).abs().max())
    scale = float(ref.abs().max())
    return err / scale


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("K,N", [(64, 64), (64, 32), (64, 128), (32, 64), (128, 64), (64, 256), (8, 16)])
def test_tc_gemm_3xtf32(mode, K, N):
    rel = _run(K, N, mode, 3)
    print("mode %d K %d N %d: 3xTF32 max rel err %.3g" % (mode, K, N, rel))
    assert rel < 2e-6


@pytest.mark.parametrize("mode", [0, 1])
def test_tc_gemm_plain_tf32_is_coarser(mode):
    rel1 = _run(64, 64, mode, 1)
    rel3 = _run(64, 64, mode, 3)
    print("mode %d: 1xTF32 %.3g, 3xTF32 %.3g" % (mode, rel1, rel3))
    assert 1e-5 < rel1 < 5e-3 and rel3 < rel1 / 50
