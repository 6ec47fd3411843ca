"""cfg2 rollout (B=200, 2000 candidates, 34 steps) timed on the legacy default stream and on a side stream:
A/B of the programmatic dependent launches of aline_rollout (ALINE_PDL=0 disables them)."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from aline_b200.attrdict import AttrDict  # noqa: E402
from aline_b200.model import Aline, Embedder, Encoder, OutputHead  # noqa: E402
from aline_b200.tasks import HiddenLocation  # noqa: E402

torch.manual_seed(123)
model = Aline(Embedder(2, 1, 32, 128, 2, "theta"), Encoder(32, 128, 4, 0.0, 3), OutputHead(2, 1, 32, 128)).cuda().eval()
task = HiddenLocation(n_query_init=int(os.environ.get("NQ", 2000)), design_scale=1)
hb = task.sample_batch(int(os.environ.get("B", 200)))
T = int(os.environ.get("T", 34))
b0 = {k: hb[k].cuda() for k in ("context_x", "context_y", "query_x", "query_y", "target_all")}


def run(stream, n=12):
    ts = []
    with torch.cuda.stream(stream):
        for _ in range(n):
            b = AttrDict(dict(b0))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model.rollout(b, T)
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:])
    return ts[len(ts) // 2]


print(json.dumps({"pdl": os.environ.get("ALINE_PDL", "1"), "default_stream_ms": run(torch.cuda.default_stream()),
                  "side_stream_ms": run(torch.cuda.Stream())}))
