"""Row f2 (training forward), host part: the differentiable torch-op composition of ``model/grad_path.py`` -- structured
attention, no [N, N] mask -- against the loss and the parameter gradients the UNMODIFIED reference produced for the same
weights, batch and (teacher-forced) sampled designs in the inner loop of train_aline.py:80-132 (fixtures
train_*.npz, tests/golden/make_golden.py gen_train)."""
import numpy as np
import pytest
import torch

from _util import load_golden, state_dict_of, train_inner_loop

pytestmark = pytest.mark.needs_grad


def run_fixture(name, device, through_forward):
    from aline_b200.attrdict import AttrDict
    from aline_b200.model import Aline, Embedder, Encoder, OutputHead, grad_path
    from aline_b200.utils.target_mask import select_targets_by_mask
    from aline_b200 import rollout as ro
    g = load_golden(name)
    sd = state_dict_of(g)
    mode = "mix" if "batch0/target_x" in g else "theta"
    dx = g["batch0/query_x"].shape[-1]
    ntok = sd["embedder.theta_tokens"].shape[0]
    model = Aline(Embedder(dx, 1, 32, 128, ntok, mode), Encoder(32, 128, 4, 0.0, 3), OutputHead(dx, 1, 32, 128))
    model.load_state_dict(sd)
    model = model.to(device).train()
    batch = AttrDict({k[7:]: torch.from_numpy(v).to(device) for k, v in g.items() if k.startswith("batch0/")})
    if "target_mask" in g:
        batch.target_mask = torch.from_numpy(g["target_mask"])
    forced = torch.from_numpy(g["idx"]).to(device)
    step = {"t": 0}

    def sampler(zt):                                   # teacher forcing: the designs the reference sampled
        i = forced[step["t"]]
        step["t"] += 1
        return i

    def update_batch(b, idx):                          # tasks/base_task.py:103-154 on plain tensors (no kernel on CPU)
        B = idx.shape[0]
        ar = torch.arange(B, device=idx.device)
        out = AttrDict(dict(b))
        for kq, kc in (("query_x", "context_x"), ("query_y", "context_y")):
            q = b[kq]
            keep = torch.ones(q.shape[:2], dtype=torch.bool, device=q.device)
            keep[ar, idx[:, 0]] = False
            out[kc] = torch.cat([b[kc], q[ar, idx[:, 0]].unsqueeze(1)], 1)
            out[kq] = q[keep].view(B, -1, q.shape[-1])
        return out

    class _Fwd:
        def forward(self, b):
            if through_forward:
                return model.forward(b)
            return grad_path.forward_torch(model, b, sampler=sampler)

    if through_forward:
        model.design_sampler = sampler
    loss, dl, pl, idx = train_inner_loop(_Fwd(), update_batch, ro.gmm_log_likelihood, select_targets_by_mask, batch,
                                         int(g["T"]), mix_n_theta=int(g["mix_n_theta"]))
    model.zero_grad()
    loss.backward()
    assert torch.equal(idx.cpu(), torch.from_numpy(g["idx"]))
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * max(1.0, abs(float(g["loss"])))
    assert abs(float(dl) - float(g["design_loss"])) < 2e-5
    worst = 0.0
    for k, p in model.named_parameters():
        ref = torch.from_numpy(g["grad/" + k])
        got = p.grad.cpu() if p.grad is not None else torch.zeros_like(ref)
        scale = max(float(ref.abs().max()), 1e-3)
        err = float((got - ref).abs().max()) / scale
        worst = max(worst, err)
        assert err < 1e-4, f"gradient of {k} differs from the reference by {err:.2e} of its scale"
    return worst


@pytest.mark.parametrize("name", ["train_location", "train_gpmix_theta"])
def test_grad_path_matches_reference_gradients(name):
    run_fixture(name, "cpu", through_forward=False)


def test_aline_forward_with_grad_refuses_cpu_parameters():
    from aline_b200._lib import AlineError
    from aline_b200.attrdict import AttrDict
    from aline_b200.model import Aline, Embedder, Encoder, OutputHead
    model = Aline(Embedder(2, 1, 32, 128, 2, "theta"), Encoder(32, 128, 4, 0.0, 3), OutputHead(2, 1, 32, 128)).train()
    b = AttrDict(context_x=torch.rand(2, 1, 2), context_y=torch.rand(2, 1, 1), query_x=torch.rand(2, 5, 2),
                 query_y=torch.rand(2, 5, 1), target_all=torch.rand(2, 2, 1))
    with pytest.raises(AlineError):
        model.forward(b)
