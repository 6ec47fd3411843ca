"""One rollout of a named secondary configuration, for ncu launch lists:
    python tools/rollout_cfg.py cfg5_d64 | cfg4_all | cfg4_data | cfg5 | cfg4_theta"""
import sys

import torch

sys.path.insert(0, ".")
from aline_b200.attrdict import AttrDict  # noqa: E402
from aline_b200.model import Aline, Embedder, Encoder, OutputHead  # noqa: E402
from aline_b200.tasks import GPTask, PsychometricTask  # noqa: E402
from aline_b200.utils.target_mask import create_target_mask  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "cfg5_d64"
torch.manual_seed(123)
if which.startswith("cfg5"):
    d = 64 if which.endswith("d64") else 32
    model = Aline(Embedder(1, 1, d, 128, 4, "theta"), Encoder(d, 128, d // 8, 0.0, 3), OutputHead(1, 1, d, 128)).cuda().eval()
    task, T, tm = PsychometricTask(n_context_init=1, n_query_init=200, design_scale=5), 30, torch.tensor([False, False, True, True])
    hb = task.sample_batch(200)
else:
    model = Aline(Embedder(2, 1, 32, 128, 3, "mix"), Encoder(32, 128, 4, 0.0, 3), OutputHead(2, 1, 32, 128)).cuda().eval()
    task, T = GPTask(dim_x=2, embedding_type="mix", n_context_init=1, n_query_init=200, n_target_theta=3, n_target_data=100,
                     design_scale=5), 50
    kind = which.split("_")[1]
    tm = {"all": create_target_mask("all", "mix", 100, 3), "data": create_target_mask("split", "mix", 100, 3, None, None, None, None, "data"),
          "theta": create_target_mask("split", "mix", 100, 3, None, None, None, None, "theta")}[kind]
    torch.set_default_device("cuda")
    hb = task.sample_batch(200)
    torch.set_default_device("cpu")
batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in hb.items()}
for _ in range(2):
    b = AttrDict(dict(batch))
    b.target_mask = tm
    out = model.rollout(b, T)
torch.cuda.synchronize()
print("ok", which)
