import sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from _util import load_golden
from aline_b200.tasks import CESTask
g = {k: torch.from_numpy(v).cuda() for k, v in load_golden("spce_ces").items()}
task = CESTask(n_context_init=1, n_query_init=1)
for t in range(2):
    ll = task.log_likelihood(g["y"][:, t].unsqueeze(0), g["x"][:, t].unsqueeze(0), g["thetas"]).squeeze(-1).cpu()
    ref = g["ll01"][t].cpu()
    th = g["thetas"].cpu()
    mu, sigma = task.response_params(g["x"][:, t].cpu().unsqueeze(0), th)
    yy = g["y"][:, t].cpu().unsqueeze(0).expand_as(mu)
    fi = torch.finfo(torch.float32)
    yc = yy.clamp(fi.tiny, 1 - fi.eps)
    z = (((yc.log() - (-yc).log1p()) - mu) / sigma).squeeze(-1)
    censored = ((yy == task.epsilon) | (yy == 1 - task.epsilon)).squeeze(-1)
    err = (ll - ref).abs()
    bad = err > 2e-2 + 1e-3 * ref.abs()
    print("t", t, "bad", int(bad.sum()), "bad&~cens", int((bad & ~censored).sum()))
    for i in torch.nonzero(bad)[:40]:
        l, b = i.tolist()
        print(f"  l={l} b={b} cens={bool(censored[l,b])} z={z[l,b].item():.5f} ll={ll[l,b].item():.6g} ref={ref[l,b].item():.6g} rho={th[l,b,0].item():.4f} logu={th[l,b,4].item():.3f}")
