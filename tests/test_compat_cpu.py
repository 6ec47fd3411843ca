"""INTEGRATION.md option (a): path shadowing through aline_b200/compat -- the reference's module paths (and hydra
`_target_` strings) resolve to the mirrors without editing the reference.  Run in a child process: the shims claim the
top-level names `model`, `loss`, `tasks`, `utils`, `distributions`."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import importlib, sys
sys.path.insert(0, sys.argv[1])
import model.base, model.embedder, model.encoder, model.head
import loss.eig, tasks.location_finding, tasks.ces, tasks.psychometric, tasks.gaussian_process, tasks.base_task
import utils.eval, utils.misc, utils.target_mask, distributions.censored_sigmoid_normal
import aline_b200
assert model.base.Aline is aline_b200.model.base.Aline
assert model.embedder.Embedder is aline_b200.model.embedder.Embedder
assert model.encoder.Encoder is aline_b200.model.encoder.Encoder
assert model.head.OutputHead is aline_b200.model.head.OutputHead
assert loss.eig.EIGStepLoss is aline_b200.loss.eig.EIGStepLoss and loss.eig.PCELoss is aline_b200.loss.eig.PCELoss
assert tasks.ces.CESTask is aline_b200.tasks.ces.CESTask
assert utils.eval.eval_boed is aline_b200.utils.eval.eval_boed
assert utils.target_mask.create_target_mask is aline_b200.utils.target_mask.create_target_mask
assert distributions.censored_sigmoid_normal.CensoredSigmoidNormal is aline_b200.distributions.CensoredSigmoidNormal
# a hydra-style `_target_` string resolves through the shim
mod, _, name = "model.embedder.Embedder".rpartition(".")
cls = getattr(importlib.import_module(mod), name)
m = model.base.Aline(cls(2, 1, 32, 128, 2, "theta"), model.encoder.Encoder(32, 128, 4, 0.0, 3), model.head.OutputHead(2, 1, 32, 128))
assert "embedder.theta_tokens" in m.state_dict()
print("ok")
'''


def test_path_shadowing_resolves_reference_module_paths():
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    out = subprocess.run([sys.executable, "-c", CHILD, os.path.join(ROOT, "aline_b200", "compat")], capture_output=True,
                         text=True, cwd="/tmp", env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
