"""Row f4: reference checkpoints (.pth state dicts, ckpt.tar) and hydra run directories load into the aline_b200
modules -- same state-dict keys, same `_target_` nodes (config/{embedder,encoder,head}/*.yaml)."""
import os

import pytest
import torch

from _util import load_golden, state_dict_of

# what hydra writes to <run>/.hydra/config.yaml for `task=psychometric` (composed, interpolations unresolved);
# values from the reference's config/train.yaml, config/{embedder,encoder,head}/*.yaml, config/task/psychometric.yaml
RUN_CONFIG = """
seed: 123
T: 30
time_token: false
file_name: aae_${task.name}.pth
output_dir: ./outputs
encoder:
  _target_: model.encoder.Encoder
  dim_embedding: %(d)d
  dim_feedforward: 128
  n_head: %(h)d
  dropout: 0.0
  num_layers: 3
embedder:
  _target_: model.embedder.Embedder
  dim_x: ${task.dim_x}
  dim_y: ${task.dim_y}
  dim_embedding: ${encoder.dim_embedding}
  dim_feedforward: ${encoder.dim_feedforward}
  n_target_theta: ${task.n_target_theta}
  embedding_type: ${task.embedding_type}
head:
  _target_: model.head.OutputHead
  dim_x: ${task.dim_x}
  dim_y: ${task.dim_y}
  dim_embedding: ${encoder.dim_embedding}
  dim_feedforward: ${encoder.dim_feedforward}
  num_components: 10
  single_head: false
  std_min: 1e-4
  value_head: false
  time_token: ${time_token}
task:
  _target_: tasks.psychometric.PsychometricTask
  name: Psychometric
  dim_x: 1
  dim_y: 1
  embedding_type: theta
  mask_type: [predefined]
  predefined_masks: [[false, false, true, true], [true, true, false, false]]
  n_context_init: 1
  n_query_init: 200
  n_target_data: 0
  n_target_theta: 4
  design_scale: 5
wandb:
  run_name: ${task.name}-${task.dim_x}D-${now:%%Y-%%m-%%d_%%H-%%M}
"""


def _run_dir(tmp_path, d, h):
    os.makedirs(tmp_path / ".hydra")
    (tmp_path / ".hydra" / "config.yaml").write_text(RUN_CONFIG % dict(d=d, h=h))
    return str(tmp_path)


@pytest.mark.parametrize("fixture,d,h", [("rollout_psychometric_a", 32, 4), ("rollout_psychometric_d64", 64, 8)])
def test_load_config_and_model(tmp_path, fixture, d, h):
    from aline_b200.utils.misc import load_config, load_config_and_model
    sd = state_dict_of(load_golden(fixture))               # the reference model's own state dict
    run = _run_dir(tmp_path, d, h)
    torch.save(sd, os.path.join(run, "aline.pth"))
    torch.save({"model": sd, "epoch": 7, "optimizer": {}, "scheduler": {}}, os.path.join(run, "ckpt.tar"))
    cfg = load_config(run)
    assert cfg.embedder.dim_x == 1 and cfg.embedder.dim_embedding == d and cfg.head.std_min == 1e-4
    assert cfg.file_name == "aae_Psychometric.pth" and cfg.head.time_token is False
    assert cfg.wandb.run_name.startswith("Psychometric-1D-${now:")        # resolver calls are left alone
    for file_name, load_type in (("aline.pth", "pth"), ("ckpt.tar", "ckpt"), ("aline.pth", "ckpt")):
        cfg, model = load_config_and_model(run, file_name=file_name, load_type=load_type, device="cpu")
        got = model.state_dict()
        assert set(got) == set(sd)
        assert all(torch.equal(got[k], sd[k]) for k in sd)
        assert model.encoder.n_head == h and model.head.target_head.std_min == pytest.approx(1e-4)
        assert model.embedder.embedding_type == "theta" and not model.head.time_token
    with pytest.raises(RuntimeError):
        load_config_and_model(run, file_name="aline.pth", load_type="zip", device="cpu")
    with pytest.raises(FileNotFoundError):
        load_config_and_model(run, file_name="missing.pth", device="cpu")
    with pytest.raises(FileNotFoundError):
        load_config_and_model(str(tmp_path / "nowhere"), device="cpu")


def test_state_dict_and_checkpoint_helpers(tmp_path):
    from aline_b200.model import Aline, Embedder, Encoder, OutputHead
    from aline_b200.utils.misc import load_checkpoint_weights, load_state_dict, save_state_dict
    sd = state_dict_of(load_golden("rollout_location_tt"))   # time-token model: predictor.0.weight is [128, 33]
    model = Aline(Embedder(2, 1, 32, 128, 2, "theta"), Encoder(32, 128, 4, 0.0, 3),
                  OutputHead(2, 1, 32, 128, time_token=True))
    model.load_state_dict(sd)
    path = save_state_dict(model, str(tmp_path), "aae_location.pth")
    assert path.endswith(os.path.join("model", "aae_location.pth"))
    fresh = Aline(Embedder(2, 1, 32, 128, 2, "theta"), Encoder(32, 128, 4, 0.0, 3),
                  OutputHead(2, 1, 32, 128, time_token=True))
    load_state_dict(fresh, str(tmp_path), "aae_location.pth")
    assert all(torch.equal(fresh.state_dict()[k], sd[k]) for k in sd)
    torch.save({"model": sd, "epoch": 11}, tmp_path / "ckpt.tar")
    other = Aline(Embedder(2, 1, 32, 128, 2, "theta"), Encoder(32, 128, 4, 0.0, 3),
                  OutputHead(2, 1, 32, 128, time_token=True))
    assert load_checkpoint_weights(other, str(tmp_path / "ckpt.tar")) == 11
    assert all(torch.equal(other.state_dict()[k], sd[k]) for k in sd)
    with pytest.raises(FileNotFoundError):
        load_checkpoint_weights(other, str(tmp_path / "none.tar"))
    mismatched = Aline(Embedder(2, 1, 32, 128, 2, "theta"), Encoder(32, 128, 4, 0.0, 3), OutputHead(2, 1, 32, 128))
    with pytest.raises(RuntimeError):                         # no time token: shape mismatch is an error, like torch
        mismatched.load_state_dict(sd)
