import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from bench_configs import model_for, cuda_batch, timeit
from aline_b200.attrdict import AttrDict
from aline_b200.tasks import HiddenLocation
task = HiddenLocation(n_query_init=2000, design_scale=1)
model = model_for(2, 2, "theta")
hb = cuda_batch(task, 200)
def roll():
    model.rollout(AttrDict(dict(hb)), 34)
for _ in range(3):
    print("rollout_ms", timeit(roll, warm=3, it=20))
