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
    top = ref.max(0, keepdim=True).values
    d = (ll - ref).abs()
    rel = d / ref.abs().clamp_min(1.0)
    print("t", t, "max abs", d.max().item(), "max rel", rel.max().item())
    idx = torch.topk(rel.flatten(), 8).indices
    for i in idx:
        l, b = divmod(i.item(), ref.shape[1])
        th = g["thetas"][l, b].cpu().numpy()
        print(f"  l={l} b={b} ll={ll[l,b].item():.6g} ref={ref[l,b].item():.6g} top-ref={top[0,b].item()-ref[l,b].item():.4g} rho={th[0]:.4f} y={g['y'][b,t,0].item():.8g}")
    near = ref > top - 30
    print("  near count", near.sum().item(), "max abs near", d[near].max().item() if near.any() else None)
