"""Host enqueue time vs device time of back-to-back cfg2 rollouts (is the rollout launch-bound?)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from aline_b200 import _lib, rollout as ro  # noqa: E402
from aline_b200.attrdict import AttrDict  # noqa: E402
from aline_b200.model import Aline, Embedder, Encoder, OutputHead  # noqa: E402
from aline_b200.tasks import HiddenLocation  # noqa: E402

torch.manual_seed(123)
model = Aline(Embedder(2, 1, 32, 128, 2, "theta"), Encoder(32, 128, 4, 0.0, 3), OutputHead(2, 1, 32, 128)).cuda().eval()
model.precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
task = HiddenLocation(n_query_init=2000, design_scale=1)
hb = task.sample_batch(200)
res = {k: hb[k].cuda() for k in ("context_x", "context_y", "query_x", "query_y", "target_all")}
lib = _lib.lib()
c_time = [0.0]
orig = lib.aline_rollout


def timed_c(*a):
    t0 = time.perf_counter()
    r = orig(*a)
    c_time[0] += time.perf_counter() - t0
    return r


class Proxy:
    def __getattr__(self, k):
        return timed_c if k == "aline_rollout" else getattr(lib, k)


_lib_lib = _lib.lib
_lib.lib = lambda: Proxy()
for _ in range(3):
    model.rollout(AttrDict(dict(res)), 34)
torch.cuda.synchronize()
N = 6
c_time[0] = 0.0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
host = []
e0.record()
for _ in range(N):
    t0 = time.perf_counter()
    model.rollout(AttrDict(dict(res)), 34)
    host.append((time.perf_counter() - t0) * 1e3)
e1.record()
torch.cuda.synchronize()
print("device ms / rollout", e0.elapsed_time(e1) / N)
print("host enqueue ms / rollout", [round(h, 2) for h in host])
print("of which inside the aline_rollout C call, ms / rollout", c_time[0] * 1e3 / N)
