"""Multi-GPU check on real NCCL (launch with torchrun, one rank per GPU):
  1. L-sharded bound: compute_EIG_from_history(shard=True) with the same explicit thetas on every rank == the
     unsharded bound computed on rank 0's GPU (1e-5), torch-generator and device-prior (Philox) variants;
  2. rollout-sharded eval_boed: outer batches dealt over ranks, bounds all-gathered -> every rank returns the same
     statistics, with finite values.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_multi_gpu.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from aline_b200.model import Aline, Embedder, Encoder, OutputHead  # noqa: E402
from aline_b200.tasks import HiddenLocation  # noqa: E402
from aline_b200.utils.eval import compute_EIG_from_history, eval_boed  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
task = HiddenLocation(n_query_init=256, design_scale=1)
torch.manual_seed(5)                                   # same histories and draws on every rank
B, T, L = 16, 20, 50_001
theta0 = torch.rand(B, 1, 2)
x = torch.rand(B, T, 2)
y = torch.log(0.1 + 1.0 / (1e-4 + ((x - theta0) ** 2).sum(-1, keepdim=True))) + 0.5 * torch.randn(B, T, 1)
thetas = torch.rand(L, B, 1, 2)
theta0, x, y, thetas = theta0.to(dev), x.to(dev), y.to(dev), thetas.to(dev)
p1, n1 = compute_EIG_from_history(task, theta0, x, y, L=L, batch_size=B, stepwise=True, thetas=thetas, shard=False)
p2, n2 = compute_EIG_from_history(task, theta0, x, y, L=L, batch_size=B, stepwise=True, thetas=thetas, shard=True)
ok_shard = torch.allclose(p1, p2, rtol=1e-5, atol=1e-5) and torch.allclose(n1, n2, rtol=1e-5, atol=1e-5)
p3, _ = compute_EIG_from_history(task, theta0, x, y, L=L, batch_size=B, stepwise=True, prior="device", seed=77, shard=False)
p4, _ = compute_EIG_from_history(task, theta0, x, y, L=L, batch_size=B, stepwise=True, prior="device", seed=77, shard=True)
ok_dev = torch.allclose(p3, p4, rtol=1e-5, atol=1e-5)
torch.manual_seed(123)
model = Aline(Embedder(2, 1, 32, 128, 2, "theta"), Encoder(32, 128, 4, 0.0, 3), OutputHead(2, 1, 32, 128)).to(dev).eval()
torch.manual_seed(1000 + rank)                         # different rollouts per rank
with torch.device(dev):
    res = eval_boed(model, task, T=10, L=20_000, M=6 * 8, batch_size=8, stepwise=True, verbose=False)
# rank-count invariance: the resident evaluation (device-side Philox batches and contrastive draws keyed by GLOBAL
# rollout / row indices) must give the same statistics for any number of ranks -- compare the printed value across runs
res_dev = eval_boed(model, task, T=10, L=20_000, M=6 * 8, batch_size=8, stepwise=True, verbose=False, prior="device", seed=11)
stats = torch.stack([res.pce_mean, res.nmc_mean]).to(dev)
ref = stats.clone()
dist.broadcast(ref, 0)
ok_boed = bool(torch.isfinite(stats).all()) and torch.equal(stats, ref) and res.pce_mean.shape == (11,)
flags = torch.tensor([ok_shard, ok_dev, ok_boed], dtype=torch.int32, device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world={world} sharded_bound_ok={bool(flags[0])} device_prior_sharded_ok={bool(flags[1])} eval_boed_ok={bool(flags[2])} "
          f"pce_T={res.pce_mean[-1].item():.4f} device_prior_pce_T={res_dev.pce_mean[-1].item():.6f} "
          f"device_prior_nmc_T={res_dev.nmc_mean[-1].item():.6f}")
dist.destroy_process_group()
sys.exit(0 if bool(flags.min()) else 1)
