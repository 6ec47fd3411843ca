"""Small end-to-end run for compute-sanitizer memcheck: forward (bf16 + fp32), resident rollout with the fused select
(16- and 32-key kernels, multi-rollout units), uncertainty rollout, step / fused (one-pass, last-only, multi-pass) /
device-prior sPCE, CES bound, d = 64 / 8-head model (warp context kernel + query_tc5), GP draw."""
import sys

import torch

sys.path.insert(0, ".")
from aline_b200.attrdict import AttrDict  # noqa: E402
from aline_b200.loss.eig import EIGStepLoss  # noqa: E402
from aline_b200.model import Aline, Embedder, Encoder, OutputHead  # noqa: E402
from aline_b200.tasks import CESTask, GPTask, HiddenLocation, PsychometricTask  # noqa: E402
from aline_b200.utils.eval import compute_EIG_from_history  # noqa: E402

torch.manual_seed(0)
dev = torch.device("cuda", 0)
model = Aline(Embedder(2, 1, 32, 128, 2, "theta"), Encoder(32, 128, 4, 0.0, 3), OutputHead(2, 1, 32, 128)).to(dev).eval()
task = HiddenLocation(n_query_init=300, design_scale=1)
hb = task.sample_batch(5)
for prec in ("bf16", "fp32"):
    model.precision = prec
    b = AttrDict({k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in hb.items()})
    out = model(b)
    _ = out.posterior_out_query.mixture_means.sum().item()
    b = AttrDict({k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in hb.items()})
    r = model.rollout(b, 20)                      # 3..22 keys: 16- and 32-key kernels
    assert int(r.query_alive.sum()) == 5 * (300 - 20)
model.precision = "bf16"
x, y = r.context_x, r.context_y
theta0 = hb["target_all"].reshape(5, 1, 2).to(dev)
with torch.device(dev):
    p, n = compute_EIG_from_history(task, theta0, x, y, L=3000, batch_size=5, stepwise=True)
    p2, _ = compute_EIG_from_history(task, theta0, x, y, L=3000, batch_size=5, stepwise=True, prior="device", seed=3)
    p3, _ = compute_EIG_from_history(task, theta0, x, y, L=3000, batch_size=5, stepwise=False)      # last-only kernel
# histories longer than one pass of the one-pass kernel (T = 40 > 36): the multi-pass fused kernel
xl, yl = torch.rand(5, 40, 2, device=dev), torch.randn(5, 40, 1, device=dev)
with torch.device(dev):
    p4, _ = compute_EIG_from_history(task, theta0, xl, yl, L=1500, batch_size=5, stepwise=True)
# CES bound (two passes of the streaming kernel) on a device-sampled batch
from aline_b200 import prior as dprior  # noqa: E402
ctask = CESTask(n_context_init=1, n_query_init=12)
cb = dprior.sample_batch_device(ctask, 4, seed=5, device=dev)
cth = dprior.sample_theta_device(ctask, 2500, 4, seed=6, device=dev)
p5, _ = compute_EIG_from_history(ctask, cb["target_all"].reshape(4, -1), ctask.unnormalise_design(cb["query_x"]),
                                 cb["query_y"], L=2500, batch_size=4, stepwise=True, thetas=cth)
# d = 64 / 8 heads: warp-per-token context kernel (weights streamed in one-matrix segments) + query_tc5
ptask = PsychometricTask(n_context_init=1, n_query_init=140, design_scale=5)
pm64 = Aline(Embedder(1, 1, 64, 128, 4, "theta"), Encoder(64, 128, 8, 0.0, 3), OutputHead(1, 1, 64, 128)).to(dev).eval()
pb = ptask.sample_batch(6)
for prec in ("bf16", "fp32"):
    pm64.precision = prec
    b = AttrDict({k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in pb.items()})
    b.target_mask = torch.tensor([False, False, True, True])
    o64 = pm64(b)
    r64 = pm64.rollout(b, 20)
    assert int(r64.query_alive.sum()) == 6 * (140 - 20)
B = 8
th = torch.rand(2001, B, 1, 2, device=dev)
xs, ys = torch.rand(B, 3, 2, device=dev), torch.randn(B, 3, 1, device=dev)
crit = EIGStepLoss(2000, B, task.log_likelihood, reduction="none")
for t in range(3):
    pl, nl = crit(ys[:, t], xs[:, t], th)
gp = GPTask(dim_x=2, embedding_type="mix", n_context_init=1, n_query_init=30, n_target_theta=3, n_target_data=10, design_scale=5)
with torch.device(dev):
    gb = gp.sample_batch(3)
gm = Aline(Embedder(2, 1, 32, 128, 3, "mix"), Encoder(32, 128, 4, 0.0, 3), OutputHead(2, 1, 32, 128)).to(dev).eval()
tm = torch.tensor([False] * 10 + [True] * 3)
gb.target_mask = tm
r2 = gm.rollout(AttrDict(dict(gb)), 6)
gb2 = AttrDict(dict(gb))
r3 = gm.rollout(gb2, 4, acquisition="uncertainty_sampling")
torch.cuda.synchronize()
print("sanitize_small ok", float(p.mean()), float(p2.mean()), float(p3.mean()), float(p4.mean()), float(p5.mean()), float(pl.mean()))
