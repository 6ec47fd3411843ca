import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from _util import load_golden, rel_err
from aline_b200.tasks import CESTask
from aline_b200.utils.eval import compute_EIG_from_history
g = load_golden("spce_ces_large")
L, seed = int(g["L"]), int(g["seed"]); B, T = g["x"].shape[:2]
task = CESTask(n_context_init=1, n_query_init=1)
torch.manual_seed(seed); thetas = task.sample_theta((L, B))
c = {k: torch.from_numpy(g[k]).cuda() for k in ("theta_0", "x", "y", "pce", "nmc")}
pce, nmc = compute_EIG_from_history(task, c["theta_0"], c["x"], c["y"], L=L, batch_size=B, stepwise=True, thetas=thetas.cuda())
d = (pce - c["pce"]).cpu(); dn = (nmc - c["nmc"]).cpu()
print("pce rel", rel_err(pce.cpu(), c["pce"].cpu()), "max abs", float(d.abs().max()), "mean signed", float(d.mean()), "nmc rel", rel_err(nmc.cpu(), c["nmc"].cpu()), "mean signed", float(dn.mean()))
# float64 reference of the bound on the same draws (oracle arithmetic in double)
