"""Row f2 on the GPU: ``Aline.forward`` under autograd (model.train(), the inner loop of train_aline.py:80-132) against
the reference's loss and parameter gradients, and the fused Philox Categorical sample (model/head.py:350-354)."""
import pytest
import torch

from test_train_cpu import run_fixture

pytestmark = [pytest.mark.gpu, pytest.mark.needs_grad]


@pytest.mark.parametrize("name", ["train_location", "train_gpmix_theta"])
def test_train_inner_loop_matches_reference_gradients(name):
    worst = run_fixture(name, "cuda", through_forward=True)
    assert worst < 1e-4


def test_fused_categorical_sample_statistics():
    """aline_select_sample: idx ~ Categorical(softmax(logits)) per rollout from (seed, rollout, step) Philox streams.
    Frequencies over 20000 draws match the probabilities (chi-square), log_prob = Categorical.log_prob, the draw is a
    pure function of (seed, rollout, step), and different rollouts / steps are decorrelated."""
    from aline_b200 import rollout as ro
    torch.manual_seed(0)
    nq = 37
    row = torch.randn(nq) * 1.5
    B = 20000
    logits = row.repeat(B, 1).cuda()
    idx, lp, zt = ro.select_sample(logits, seed=1234, step=3)
    p = torch.softmax(row.double(), 0)
    counts = torch.bincount(idx[:, 0].cpu(), minlength=nq).double()
    chi2 = float(((counts - B * p) ** 2 / (B * p)).sum())
    assert chi2 < 80.0, f"chi-square {chi2:.1f} for {nq - 1} degrees of freedom"          # p ~ 1e-5 tail at 80
    ref_lp = torch.log(p.float().clamp(1.1920929e-07, 1 - 1.1920929e-07))[idx[:, 0].cpu()]
    assert torch.allclose(lp.cpu(), ref_lp, rtol=1e-5, atol=1e-6)
    assert torch.allclose(zt.cpu()[0].double(), p, rtol=1e-5, atol=1e-8)
    idx2, _, _ = ro.select_sample(logits, seed=1234, step=3)
    assert torch.equal(idx, idx2)
    idx3, _, _ = ro.select_sample(logits, seed=1234, step=4)
    assert float((idx3 == idx).double().mean()) < 0.5
    # in-place append + retire with sampling (resident train-mode step): chosen candidates are live ones
    Bs, nq2 = 64, 300
    lg = torch.randn(Bs, nq2, device="cuda")
    alive = torch.ones(Bs, nq2, dtype=torch.uint8, device="cuda")
    alive[:, ::3] = 0
    from aline_b200 import _lib
    import ctypes
    idx_o = torch.empty(Bs, 1, dtype=torch.int64, device="cuda")
    orig = torch.empty(Bs, dtype=torch.int64, device="cuda")
    lpo = torch.empty(Bs, device="cuda")
    _lib.check(_lib.lib().aline_select_sample(_lib.dptr(lg), _lib.dptr(alive, torch.uint8), Bs, nq2, None, None, 0, 0, None,
                                              None, 0, 0, _lib.dptr(idx_o, torch.int64), 1, _lib.dptr(lpo), 1,
                                              _lib.dptr(orig, torch.int64), None, ctypes.c_uint64(99), 0,
                                              _lib.stream_ptr(lg.device)))
    torch.cuda.synchronize()
    assert bool((orig % 3 != 0).all()) and bool((orig >= 0).all()) and bool((orig < nq2).all())
    live_before = torch.stack([(torch.arange(nq2, device="cuda")[None, :] < orig[:, None]) & (torch.arange(nq2, device="cuda")[None, :] % 3 != 0)]).sum(-1)[0]
    assert torch.equal(idx_o[:, 0], live_before)


def test_train_mode_forward_without_grad_uses_fused_sample():
    from aline_b200.attrdict import AttrDict
    from aline_b200.model import Aline, Embedder, Encoder, OutputHead
    torch.manual_seed(5)
    model = Aline(Embedder(2, 1, 32, 128, 2, "theta"), Encoder(32, 128, 4, 0.0, 3), OutputHead(2, 1, 32, 128)).cuda().train()
    b = AttrDict(context_x=torch.rand(8, 1, 2).cuda(), context_y=torch.rand(8, 1, 1).cuda(), query_x=torch.rand(8, 50, 2).cuda(),
                 query_y=torch.rand(8, 50, 1).cuda(), target_all=torch.rand(8, 2, 1).cuda())
    with torch.no_grad():
        out = model.forward(b)
    idx, lp, zt = out.design_out.idx, out.design_out.log_prob, out.design_out.zt
    assert idx.shape == (8, 1) and idx.dtype == torch.int64 and bool((idx >= 0).all()) and bool((idx < 50).all())
    assert torch.allclose(lp, torch.log(zt.gather(1, idx)[:, 0]), rtol=1e-5, atol=1e-6)
