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


# ------------------------------------------------------------------ tensor-core cost volume vs the SIMT fp32 kernel

def _cv_inputs(B, N1, N2, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    Gab, Hab, H3 = r(B, N2, 128) * 0.7, r(B, N1, 128) * 0.7, r(B, N1, 64) * 0.5
    xyz1, xyz2 = r(B, N1, 3) * 3, r(B, N2, 3) * 3
    idx = torch.randint(0, N2, (B, N1, 16), generator=g, dtype=torch.int32)
    idxw = torch.randint(0, N2, (B, N1, 16), generator=g, dtype=torch.int32)
    return [t.cuda().contiguous() for t in (Gab, Hab, H3, xyz1, xyz2, idx, idxw)]


@pytest.fixture(scope="module")
def cv_weights():
    from oracle import tflow_port as tp  # checker-side weight generator only
    from ssf_slam_b200.model import prepare_weights
    return prepare_weights(tp.random_init_state_dict(0), torch.device("cuda:0"))


@pytest.mark.parametrize("level,B,N1,N2,n_sm", [("flow1_r", 1, 8, 40, 0), ("flow1_r", 2, 203, 150, 0), ("flow0_r", 3, 2048, 1024, 0),
                                                ("flow0_r", 2, 1024, 512, 3)])
def test_cost_volume_tc_matches_simt(cv_weights, level, B, N1, N2, n_sm):
    from ssf_slam_b200 import functional as F_
    from ssf_slam_b200 import _native as nat
    w = cv_weights[level]
    Gab, Hab, H3, xyz1, xyz2, idx, idxw = _cv_inputs(B, N1, N2, 7)
    F_.USE_TC = False
    try:
        ref = F_.cost_volume(Gab, Hab, w, H3, xyz1, xyz2, idx, idxw, 64)
    finally:
        F_.USE_TC = True
    if n_sm == 0:
        got = F_.cost_volume(Gab, Hab, w, H3, xyz1, xyz2, idx, idxw, 64)
    else:  # few CTAs -> many tiles per persistent CTA (exercises the barrier phase bookkeeping)
        got = [torch.full_like(t, float("nan")) for t in ref]
        p = nat.ptr
        nat.check(nat.lib().ssf_cost_volume_tc(p(Gab), p(Hab), p(H3), p(w["tc_blob"]), p(w["tc_par"]), p(xyz1), p(xyz2), p(idx),
                                               p(idxw), B, N1, N2, 64, p(got[0]), p(got[1]), p(got[2]), p(got[3]), n_sm,
                                               nat.stream()))
    torch.cuda.synchronize()
    for name, r_, g_ in zip(("cost_fwd", "cost_fwd_cm", "gw", "Cw"), ref, got):
        scale = max(1.0, float(r_.abs().max()))
        err = float((r_ - g_).abs().max())
        print(level, name, "max-abs diff %.3g (scale %.3g)" % (err, scale))
        assert err < 2e-5 * scale, (name, err, scale)
