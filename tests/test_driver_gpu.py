"""The training driver end to end on the GPU: R/train.py's command line, run-directory rules,
checkpoint + validation cadence and resume, on a small config with the synthetic data source."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from test_driver_cpu import write_config  # noqa: E402


def test_train_driver_runs_saves_validates_and_resumes(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from plbert_b200 import train as T
    cfg_path = tmp_path / "config.yml"
    write_config(cfg_path, tmp_path / "runs", synthetic_batches_per_epoch=4)
    args = T.parse_args(["--config_path", str(cfg_path), "--run_name", "r1", "--no_wandb"])
    step, epoch = T.train(args)
    assert (step, epoch) == (6, 2)                       # 4 batches per epoch, num_steps 6
    run = tmp_path / "runs" / "r1"
    assert {"config.yml", "metrics.jsonl", "step_3.t7", "step_6.t7"} <= set(os.listdir(run))
    rows = [json.loads(x) for x in open(run / "metrics.jsonl")]
    train_rows = [r for r in rows if "phoneme_loss" in r]
    val_rows = [r for r in rows if "val_phoneme_loss" in r]
    assert len(train_rows) == 6 and [r["step"] for r in val_rows] == [0, 3, 6]      # R/train.py:342,366-371
    assert all("phoneme_loss_avg" in r for r in train_rows[1:]) and "phoneme_loss_avg" not in train_rows[0]
    assert all(0 < r["phoneme_loss"] < 12 for r in train_rows) and all(0 < r["val_phoneme_loss"] < 12 for r in val_rows)
    assert val_rows[-1]["val_phoneme_loss"] < val_rows[0]["val_phoneme_loss"]         # six steps already help
    ck = torch.load(run / "step_6.t7", map_location="cpu", weights_only=False)
    assert ck["step"] == 6 and ck["epoch"] == 2 and all(k.startswith("module.") for k in ck["net"])
    assert len(ck["optimizer"]["state"]) == len(ck["net"]) - 2                          # no state for the pooler
    # validation must not touch weights or optimizer state
    model, opt, cur = T.initialize_model(T.setup_config_and_directories(args, str(cfg_path))[0], str(run), True, "cuda:0")
    assert cur == 6
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    _, val_dl = T.setup_dataset_and_dataloader(ck_cfg(cfg_path), "cuda:0")
    v1 = T.validate(model, val_dl)
    v2 = T.validate(model, val_dl)
    assert v1 == v2 and abs(v1 - val_rows[-1]["val_phoneme_loss"]) < 1e-3 * v1         # same weights, same batches
    assert all(torch.equal(p.detach(), before[n]) for n, p in model.named_parameters())
    # resume: the run directory's own config + newest checkpoint; the reference then runs one more step and stops
    step2, _ = T.train(args)
    assert step2 == 7


def test_train_driver_b200_params_block(tmp_path):
    """The optional ``b200_params`` block of the YAML: two-head training with the token loss over every valid position,
    reference-style ``.pth`` checkpoints."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from plbert_b200 import train as T
    cfg_path = tmp_path / "config.yml"
    write_config(cfg_path, tmp_path / "runs", synthetic_batches_per_epoch=4, num_steps=2, save_interval=2,
                 b200=dict(use_token_ids=True, num_tokens=300, token_rows="all_valid", checkpoint_extension="pth"))
    step, _ = T.train(T.parse_args(["--config_path", str(cfg_path), "--run_name", "mt", "--no_wandb"]))
    assert step == 2
    run = tmp_path / "runs" / "mt"
    ck = torch.load(run / "step_2.pth", map_location="cpu", weights_only=False)
    assert ck["net"]["module.token_predictor.weight"].shape == (300, 128)
    rows = [json.loads(x) for x in open(run / "metrics.jsonl")]
    losses = [r["phoneme_loss"] for r in rows if "phoneme_loss" in r]
    assert len(losses) == 2 and all(8 < x < 14 for x in losses)      # ln(188) + ln(300) = 10.9 at initialisation


def ck_cfg(path):
    import yaml
    with open(path) as f:
        return yaml.safe_load(f)
