"""One cfg2 rollout (B=200, 2000 candidates, 34 steps) after a warm-up: the command profiled for profiles/*launches*."""
import sys
import torch
sys.path.insert(0, ".")
from aline_b200.attrdict import AttrDict
from aline_b200.model import Aline, Embedder, Encoder, OutputHead
from aline_b200.tasks import HiddenLocation
torch.manual_seed(123)
model = Aline(Embedder(2, 1, 32, 128, 2, "theta"), Encoder(32, 128, 4, 0.0, 3), OutputHead(2, 1, 32, 128)).cuda().eval()
model.precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
task = HiddenLocation(n_query_init=2000, design_scale=1)
hb = task.sample_batch(200)
import time
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    b = AttrDict({k: hb[k].cuda() for k in ("context_x", "context_y", "query_x", "query_y", "target_all")})
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    out = model.rollout(b, 34)
    t1 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    print("rollout ms", e0.elapsed_time(e1), "host enqueue ms", (t1 - t0) * 1e3)
