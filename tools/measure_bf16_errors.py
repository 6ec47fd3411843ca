import sys, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from _util import load_golden, state_dict_of, abs_err, rel_err
from test_forward_gpu import build_model, attr_batch, step_batch, mode_of, ROLLOUTS
torch.set_grad_enabled(False)
for name in ROLLOUTS:
    g = load_golden(name); sd = state_dict_of(g)
    model = build_model(sd, mode_of(g), precision="bf16")
    mx_row, mx_pq, mx_lp = 0, {}, 0
    for t in range(int(g["n_steps"])):
        b = attr_batch(step_batch(g, t)); pred = model.forward(b); pre = f"step{t}/"
        zt, zr = pred.design_out.zt.cpu().double(), torch.from_numpy(g[pre + "zt"]).double()
        mx_row = max(mx_row, float((zt.log() - zr.log()).abs().max()))
        mx_lp = max(mx_lp, rel_err(pred.design_out.log_prob.cpu(), g[pre + "log_prob"]))
        for k in ("mixture_means", "mixture_stds", "mixture_weights"):
            mx_pq[k] = max(mx_pq.get(k, 0), abs_err(pred.posterior_out_query[k].cpu(), g[pre + "postq/" + k]))
    lg = torch.from_numpy(g["step0/logits"]).double()
    print(f"{name:28s} row_err {mx_row:.2e} logp_rel {mx_lp:.2e} spread {float(lg.max()-lg.min()):.3f} postq", {k: f"{v:.1e}" for k, v in mx_pq.items()})
