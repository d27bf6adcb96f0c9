"""Host-side logic of the "next" rows of SURVEY.md §8f that needs no GPU: run-directory / resume rules and batch
sharding of the training driver (f1), the look-ahead feeder's bookkeeping (f2), the checkpoint -> Hugging Face
converter and the StyleTTS loading recipe (f3)."""
import json
import os

import pytest
import torch
import yaml

from plbert_b200 import train as T
from plbert_b200.batch import BatchFeeder, synthetic_collate_batch

TINY = dict(hidden_size=128, num_attention_heads=2, intermediate_size=256, max_position_embeddings=64, num_hidden_layers=2)


def write_config(path, out_dir, b200=None, **training):
    cfg = {"training_params": {**dict(output_dir=str(out_dir), mixed_precision="fp16", batch_size=4, save_interval=3, log_interval=2,
                                      num_steps=6, learning_rate=7e-5, training_dataset="synthetic", split="x"), **training},
           "dataset_params": dict(word_separator=87, max_seq_length=64, word_pred_prob=0.3, phoneme_mask_prob=0.8, replace_prob=0.1),
           "model_params": dict(pretrained_model="", dropout=0.1, **TINY)}
    if b200 is not None:
        cfg["b200_params"] = b200
    with open(path, "w") as f:
        yaml.dump(cfg, f)
    return cfg


def test_run_directory_rules(tmp_path):
    """R/train.py:164-204: fresh dir -> created + config copied; dir with config -> resume from THAT config; dir without
    config -> step_* files purged, fresh start."""
    cfg_path = tmp_path / "config.yml"
    write_config(cfg_path, tmp_path / "runs")
    args = {"run_name": "a"}
    cfg, log_dir, resuming = T.setup_config_and_directories(args, str(cfg_path))
    assert not resuming and os.path.exists(os.path.join(log_dir, "config.yml")) and cfg["training_params"]["num_steps"] == 6
    # edit the ORIGINAL config: a resumed run must keep using the copy in the run directory
    write_config(cfg_path, tmp_path / "runs", extra_key=1)
    cfg2, log_dir2, resuming2 = T.setup_config_and_directories(args, str(cfg_path))
    assert resuming2 and log_dir2 == log_dir and "extra_key" not in cfg2["training_params"]
    # a run directory that lost its config is purged of checkpoints and restarted
    os.remove(os.path.join(log_dir, "config.yml"))
    for n in ("step_3.t7", "step_6.pth", "notes.txt"):
        open(os.path.join(log_dir, n), "w").close()
    cfg3, _, resuming3 = T.setup_config_and_directories(args, str(cfg_path))
    assert not resuming3 and "extra_key" in cfg3["training_params"]
    assert sorted(os.listdir(log_dir)) == ["config.yml", "notes.txt"]
    assert T.find_latest_checkpoint(log_dir) == (False, 0)


def test_shard_batch_splits_contiguously_and_repads():
    batch = synthetic_collate_batch(6, 64, seed=3, variable_length=True, min_length=8, with_token_ids=True, num_tokens=99)
    tok, lab, msk, lens, midx = batch
    assert T.shard_batch(batch, 0, 1) is batch
    parts = [T.shard_batch(batch, r, 3) for r in range(3)]
    for r, (t, l, m, ln, mi) in enumerate(parts):
        lo = 2 * r
        assert ln == lens[lo:lo + 2] and mi == midx[lo:lo + 2]
        assert t.shape == l.shape == m.shape == (2, max(ln))            # padded to the slice's own maximum
        assert torch.equal(m, msk[lo:lo + 2, :max(ln)]) and torch.equal(l, lab[lo:lo + 2, :max(ln)])
        assert torch.equal(t, tok[lo:lo + 2, :max(ln)])
    with pytest.raises(ValueError):
        T.shard_batch(batch, 0, 4)
    # validation batches (drop_last=False) may be uneven: ceil/floor shares, every sample exactly once, empty share -> None
    shares = [T.shard_batch(batch, r, 4, strict=False) for r in range(4)]
    assert [len(s[3]) for s in shares] == [2, 2, 1, 1] and sum((s[3] for s in shares), []) == lens
    assert [T.shard_batch(batch, r, 8, strict=False) is None for r in range(8)] == [False] * 6 + [True] * 2


def test_synthetic_loader_is_deterministic_and_redraws_every_epoch(tmp_path):
    cfg = write_config(tmp_path / "c.yml", tmp_path, synthetic_batches_per_epoch=3)
    a, _ = T.setup_dataset_and_dataloader(cfg, "cpu")
    b, v = T.setup_dataset_and_dataloader(cfg, "cpu")
    e1, e1b, e2 = list(a), list(b), list(a)
    assert len(e1) == len(a) == 3 and len(list(v)) == 2
    assert all(torch.equal(x[1], y[1]) and x[2] == y[2] for x, y in zip(e1, e1b))
    assert not all(x[1].shape == y[1].shape and torch.equal(x[1], y[1]) for x, y in zip(e1, e2))
    labels, masked, lengths, midx = e1[0]
    assert labels.shape == masked.shape and labels.shape[0] == 4 and labels.shape[1] <= 64 and len(midx) == 4


def test_metrics_logger_writes_jsonl(tmp_path):
    lg = T.MetricsLogger(str(tmp_path), {}, enabled=True, use_wandb=False)
    lg.log({"phoneme_loss": 1.5, "epoch": 1})
    lg.log({"val_phoneme_loss": 2.0, "step": 3, "epoch": 1})
    lg.close()
    rows = [json.loads(x) for x in open(tmp_path / "metrics.jsonl")]
    assert rows == [{"phoneme_loss": 1.5, "epoch": 1}, {"val_phoneme_loss": 2.0, "step": 3, "epoch": 1}]
    T.MetricsLogger(str(tmp_path / "nowhere"), {}, enabled=False, use_wandb=False).log({"x": 1})  # non-main ranks: no file


class _RecordingModel:
    """Stands in for the CUDA model: BatchFeeder only needs stage_batch / loss_and_backward."""

    def __init__(self):
        self.events = []

    def new_stager(self):
        return "private stager"

    def stage_batch(self, b, stager=None):
        assert stager == "private stager"
        self.events.append(("stage", b))
        return ("staged", b)

    def loss_and_backward(self, db):
        self.events.append(("step", db[1]))
        return db[1]


def test_batch_feeder_stages_one_batch_ahead():
    m = _RecordingModel()
    feed = BatchFeeder(m, iter([10, 11, 12]))
    assert m.events == [("stage", 10)] and not feed.exhausted
    assert feed.step() == 10
    assert m.events[1:] == [("step", 10), ("stage", 11)]      # next batch staged right after the step was enqueued
    assert feed.step() == 11 and feed.step() == 12 and feed.exhausted
    assert [e for e in m.events if e[0] == "stage"] == [("stage", 10), ("stage", 11), ("stage", 12)]
    with pytest.raises(StopIteration):
        feed.step()


class _FakeModel(torch.nn.Module):
    """CPU stand-in with the three calls train_loop makes; the loss is the number of the batch it was computed on."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.loss_buf = torch.zeros(1)
        self.calls = []

    def new_stager(self):
        return None

    def stage_batch(self, b, stager=None):
        return b

    def loss_and_backward(self, b):
        self.loss_buf[0] = float(b)          # like the engine: ONE scalar buffer, overwritten by every step
        self.calls.append(("step", b))
        return self.loss_buf[0]

    def loss(self, b):
        self.calls.append(("val", b))
        return torch.tensor(100.0 + b)


class _FakeOpt:
    def __init__(self):
        self.steps = 0

    def step(self):
        self.steps += 1

    def state_dict(self):
        return {"state": {}, "param_groups": []}


@pytest.mark.parametrize("log_lag", [0, 1, 3])
def test_train_loop_cadence_and_lagged_logging(tmp_path, log_lag):
    """R/train.py:338-379: validation first, checkpoint + validation every save_interval steps, stop at num_steps; with a
    logging lag every step's loss is still logged, in order, with ITS value (the loss scalar is reused by the engine)."""
    model, opt = _FakeModel(), _FakeOpt()
    logger = T.MetricsLogger(str(tmp_path), {}, enabled=True, use_wandb=False)
    train_dl, val_dl = [1, 2, 3, 4], [7, 8]                       # 4 batches per epoch
    step, epoch = T.train_loop(model, opt, train_dl, val_dl, 0, num_steps=10, save_interval=3, log_interval=2,
                               log_dir=str(tmp_path), logger=logger, log_lag=log_lag)
    logger.close()
    assert (step, epoch) == (10, 3) and opt.steps == 10
    rows = [json.loads(x) for x in open(tmp_path / "metrics.jsonl")]
    train_rows = [r for r in rows if "phoneme_loss" in r]
    assert [r["phoneme_loss"] for r in train_rows] == [1, 2, 3, 4, 1, 2, 3, 4, 1, 2]
    assert [r["epoch"] for r in train_rows] == [1, 1, 1, 1, 2, 2, 2, 2, 3, 3]
    assert train_rows[1]["phoneme_loss_avg"] == 1.5 and "phoneme_loss_avg" not in train_rows[0]
    val_rows = [r for r in rows if "val_phoneme_loss" in r]
    assert [r["step"] for r in val_rows] == [0, 3, 6, 9] and all(r["val_phoneme_loss"] == 107.5 for r in val_rows)
    assert sorted(f for f in os.listdir(tmp_path) if f.startswith("step_")) == ["step_3.t7", "step_6.t7", "step_9.t7"]
    # all losses up to a checkpoint are in the log before the checkpoint's validation row
    idx_val3 = rows.index(val_rows[1])
    assert sum(1 for r in rows[:idx_val3] if "phoneme_loss" in r) == 3


# ------------------------------------------------------------------------------------------------ converter (f3)
def reference_style_checkpoint(path, two_heads=True, prefix="module.", upstream_names=False, vocab=188, vt=321):
    """A checkpoint as R/train.py:416-424 writes it, from the stock classes the reference wraps (R/model.py:5-30)."""
    from transformers import AlbertConfig, AlbertModel
    torch.manual_seed(1)
    enc = AlbertModel(AlbertConfig(vocab_size=vocab, **TINY))
    net = {"encoder." + k: v for k, v in enc.state_dict().items()}
    ph, tk = ("mask_predictor", "word_predictor") if upstream_names else ("phoneme_predictor", "token_predictor")
    net[ph + ".weight"], net[ph + ".bias"] = torch.randn(vocab, 128), torch.randn(vocab)
    if two_heads:
        net[tk + ".weight"], net[tk + ".bias"] = torch.randn(vt, 128), torch.randn(vt)
    net["encoder.embeddings.position_ids"] = torch.arange(64)[None]   # persisted by transformers < 4.31 checkpoints
    torch.save({"net": {prefix + k: v for k, v in net.items()}, "step": 42, "epoch": 3, "optimizer": None}, path)
    return enc, net


@pytest.mark.parametrize("two_heads,prefix,upstream", [(True, "module.", False), (False, "", False), (True, "module.", True)])
def test_convert_to_hf_round_trip(tmp_path, two_heads, prefix, upstream):
    from transformers import AlbertModel
    from plbert_b200 import convert_to_hf as C
    cfg_path = tmp_path / "config.yml"
    write_config(cfg_path, tmp_path / "runs")
    ck = tmp_path / "step_42.t7"
    enc, net = reference_style_checkpoint(ck, two_heads, prefix, upstream)
    out = tmp_path / "hf"
    C.main(["--checkpoint_path", str(ck), "--config_path", str(cfg_path), "--output_dir", str(out)])
    assert {"config.json", "config.yml", "pl_bert_full_model.pt", "training_metadata.txt"} <= set(os.listdir(out))
    assert "Step: 42" in open(out / "training_metadata.txt").read() and "Epoch: 3" in open(out / "training_metadata.txt").read()
    # the encoder directory loads into the stock class with the checkpoint's vocabulary (R/convert_to_hf.py:85)
    hf = AlbertModel.from_pretrained(str(out))
    assert hf.config.vocab_size == 188
    for k, v in enc.state_dict().items():
        assert torch.equal(hf.state_dict()[k], v), k
    # full model incl. heads (R/convert_to_hf.py:66-102)
    model = C.load_pl_bert_model(str(out))
    assert not model.training and (hasattr(model, "token_predictor") == two_heads)
    ph = "mask_predictor" if upstream else "phoneme_predictor"
    assert torch.equal(model.phoneme_predictor.weight, net[ph + ".weight"])
    if two_heads:
        assert model.token_predictor.out_features == 321
        with pytest.raises(ValueError):
            C.convert(str(ck), str(cfg_path), str(tmp_path / "bad"), num_tokens=999)
    # StyleTTS recipe (R/README.md:57-66): strict load of the encoder weights into a fresh AlbertModel
    from transformers import AlbertConfig
    bert = AlbertModel(AlbertConfig(vocab_size=188, **TINY))
    sd = C.styletts_encoder_state_dict(str(ck))
    sd.pop("embeddings.position_ids", None)
    bert.load_state_dict(sd)
    assert torch.equal(bert.embeddings.word_embeddings.weight, enc.embeddings.word_embeddings.weight)


def test_convert_rejects_mismatched_checkpoint(tmp_path):
    from plbert_b200 import convert_to_hf as C
    cfg_path = tmp_path / "config.yml"
    cfg = write_config(cfg_path, tmp_path / "runs")
    cfg["model_params"]["intermediate_size"] = 512     # config disagrees with the checkpoint tensors
    yaml.dump(cfg, open(cfg_path, "w"))
    ck = tmp_path / "step_1.pth"
    reference_style_checkpoint(ck)
    with pytest.raises(RuntimeError):
        C.convert(str(ck), str(cfg_path), str(tmp_path / "out"))
    torch.save({"weights": {}}, tmp_path / "junk.pt")
    with pytest.raises(KeyError):
        C.read_checkpoint(str(tmp_path / "junk.pt"))
