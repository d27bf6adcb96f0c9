"""Mirrors of the step-level functions of R/train.py, routed to the B200 path.

Same names and argument meaning as the reference so a maintainer can swap the import:
``length_to_mask`` (R/train.py:34-44), ``process_batch`` (:381-390), ``find_latest_checkpoint`` (:46-79),
``load_checkpoint`` (:81-105), ``save_checkpoint`` (:412-425) — and the training driver around them:
``parse_args`` (:27-32), ``setup_config_and_directories`` (:164-204), ``initialize_model`` (:259-286), ``validate``
(:288-305), ``train_loop`` (:338-379), ``train`` (:133-162) with the same ``--config_path --run_name`` command line, run
directory / resume rules and checkpoint + validation cadence, but without ``accelerate``::

    python -m plbert_b200.train --config_path configs/config.yml --run_name my_run                      # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 -m plbert_b200.train ...   # one box

Data parallelism is ``ddp.GradientSynchronizer`` (one process per GPU, the global batch split across ranks like
``Accelerator(split_batches=True)``, R/train.py:220); mixed precision is the engine's bf16/fp32-master scheme whatever
``mixed_precision`` says; wandb is used when importable and not disabled, a JSON-lines file in the run directory
otherwise.  The dataset is whatever yields the reference's collate tuples: ``training_dataset: synthetic`` selects the
built-in generator, anything else goes through ``plbert_b200.data`` — the reference's masking + collate in background
worker processes.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import time
from collections import deque
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from . import _abi
from ._abi import call, ptr
from .model import normalize_state_dict_keys


def length_to_mask(lengths: torch.Tensor) -> torch.Tensor:
    """R/train.py:34-44 — bool [B, max_len], True where ``position + 1 > length`` (padding).

    ``lengths`` on a CUDA device: computed by ``plb_length_to_mask`` (bit-exact integer kernel).  The reference
    calls this on a CPU tensor (R/train.py:384); a CPU input is a host-side integer comparison, not a compute
    fallback, and is kept so existing callers keep working."""
    if lengths.dim() != 1 or lengths.numel() == 0:
        raise ValueError("lengths must be a non-empty 1-D tensor")
    B = lengths.size(0)
    max_len = int(lengths.max().item())
    if lengths.device.type == "cuda":
        li = lengths.to(torch.int32).contiguous()
        valid = torch.empty((B, max_len), dtype=torch.int32, device=lengths.device)
        call("plb_length_to_mask", ptr(li), B, max_len, ptr(valid), torch.cuda.current_stream().cuda_stream)
        return valid == 0
    positions = torch.arange(max_len).expand(B, max_len)
    return (positions + 1) > lengths.unsqueeze(1)


def process_batch(model, batch, criterion=None, accelerator=None) -> torch.Tensor:
    """R/train.py:381-390 — forward + loss for one collate batch.  On the B200 path the backward pass is fused
    with it (``model.loss_and_backward``): the returned loss is a detached device scalar and the gradients are
    already in ``p.grad``, so the caller's ``accelerator.backward(loss)`` / ``loss.backward()`` must be dropped
    (see INTEGRATION.md).  ``criterion`` / ``accelerator`` are accepted for signature compatibility."""
    return model.loss_and_backward(batch)


def find_latest_checkpoint(log_dir: str) -> Tuple[bool, int]:
    """R/train.py:46-79 — newest ``step_<int>.<ext>`` in ``log_dir`` (extension-agnostic)."""
    found, last = False, 0
    try:
        steps = []
        for f in os.listdir(log_dir):
            if f.startswith("step_") and os.path.isfile(os.path.join(log_dir, f)):
                try:
                    steps.append(int(f.split("_")[-1].split(".")[0]))
                except ValueError:
                    continue
        if steps:
            found, last = True, max(steps)
    except Exception as e:  # same forgiving behaviour as the reference
        print(f"Error finding checkpoints: {e}")
        found, last = False, 0
    return found, last


def save_checkpoint(model, optimizer, current_step: int, log_dir: str, current_epoch: int = 0, *,
                    extension: str = "t7", module_prefix: bool = True) -> str:
    """R/train.py:412-425 — ``{'net', 'step', 'epoch', 'optimizer'}`` pickled with torch.save.

    ``extension`` defaults to ``t7`` (north star); the reference writes ``.pth`` (R/train.py:423) and reads any.
    ``module_prefix`` emits the DDP ``module.`` key prefix the StyleTTS loader slices off unconditionally
    (R/README.md:62) and that R/train.py:98 / R/convert_to_hf.py:40 strip when present."""
    net = {("module." + k if module_prefix else k): v.detach().to("cpu").clone() for k, v in model.state_dict().items()}
    opt = optimizer.state_dict()
    state = {"net": net, "step": current_step, "epoch": current_epoch, "optimizer": opt}
    os.makedirs(log_dir, exist_ok=True)
    path = os.path.join(log_dir, f"step_{current_step}.{extension}")
    torch.save(state, path)
    return path


def load_checkpoint(model, optimizer, checkpoint_path: str, accelerator=None, device: Optional[torch.device] = None):
    """R/train.py:81-105 — strict=False weight load (``module.`` stripped, upstream head names aliased) +
    optimizer state."""
    map_location = device or (accelerator.device if accelerator is not None else "cpu")
    checkpoint = torch.load(checkpoint_path, map_location=map_location, weights_only=False)
    model.load_state_dict(normalize_state_dict_keys(checkpoint["net"]), strict=False)
    if optimizer is not None and checkpoint.get("optimizer") is not None:
        optimizer.load_state_dict(checkpoint["optimizer"])
    return model, optimizer


# ================================================================================================
# training driver
# ================================================================================================
def parse_args(argv=None) -> Dict[str, object]:
    """R/train.py:27-32 plus the switches a machine without the dataset / wandb needs."""
    ap = argparse.ArgumentParser(description="Train phoneme-level BERT model (B200 path)")
    ap.add_argument("--config_path", type=str, default="configs/config.yml", help="Path to config file")
    ap.add_argument("--run_name", type=str, default="default", help="Name of the run for organizing checkpoints")
    ap.add_argument("--use_token_ids", action="store_true", help="train MultiTaskModel (phoneme + token heads)")
    ap.add_argument("--num_tokens", type=int, default=None, help="token-head vocabulary when --use_token_ids")
    ap.add_argument("--no_wandb", action="store_true", help="log to <run dir>/metrics.jsonl even if wandb is installed")
    ap.add_argument("--checkpoint_extension", type=str, default="t7")
    return vars(ap.parse_args(argv))


def setup_config_and_directories(args: Dict[str, object], config_path: str):
    """R/train.py:164-204.  An existing run directory WITH a copy of the config resumes from that copy; an existing
    directory without one is purged of ``step_*`` files and restarted; otherwise the directory is created."""
    import yaml
    with open(config_path) as f:
        original = yaml.safe_load(f)
    log_dir = os.path.join(original["training_params"]["output_dir"], str(args["run_name"]))
    kept = os.path.join(log_dir, os.path.basename(config_path))
    if os.path.isdir(log_dir) and os.path.exists(kept):
        with open(kept) as f:
            return yaml.safe_load(f), log_dir, True
    if os.path.isdir(log_dir):
        for name in os.listdir(log_dir):
            if name.startswith("step_"):
                os.remove(os.path.join(log_dir, name))
    os.makedirs(log_dir, exist_ok=True)
    shutil.copy(config_path, kept)
    return original, log_dir, False


def shard_batch(batch: Sequence, rank: int, world: int, strict: bool = True):
    """This rank's contiguous slice of a global collate batch (``Accelerator(split_batches=True)``, R/train.py:220),
    re-padded to the slice's own longest sequence (the collate contract: zero padding up to the batch maximum).

    ``strict`` (training batches): the batch size must be a multiple of the number of processes, as ``split_batches=True``
    demands.  ``strict=False`` (validation: ``drop_last=False`` leaves a last batch of any size, R/dataloader.py:264-273):
    ranks get ceil/floor shares; a rank whose share is empty gets ``None``."""
    if world == 1:
        return batch
    *tensors, lengths, midx = batch
    B = len(lengths)
    if B % world and strict:
        raise ValueError(f"batch size {B} is not divisible by the number of processes {world} (split_batches=True)")
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    if hi == lo:
        return None
    lens = [int(x) for x in lengths[lo:hi]]
    width = max(lens)
    return tuple(t[lo:hi, :width].contiguous() for t in tensors) + (lens, list(midx[lo:hi]))


class SyntheticLoader:
    """Epoch-style iterable of synthetic collate batches (``training_dataset: synthetic``): ``num_batches`` batches per
    epoch, re-drawn every epoch like the reference's on-the-fly masking (R/dataloader.py:59-126)."""

    def __init__(self, batch_size: int, dataset_params: Dict[str, object], num_batches: int, seed: int, use_token_ids: bool,
                 num_tokens: int, variable_length: bool = True, redraw: bool = True):
        self.batch_size, self.num_batches, self.seed, self.redraw = batch_size, num_batches, seed, redraw
        self.kw = dict(max_seq_length=int(dataset_params.get("max_seq_length", 512)), with_token_ids=use_token_ids,
                       num_tokens=num_tokens, variable_length=variable_length,
                       word_pred_prob=float(dataset_params.get("word_pred_prob", 0.15)),
                       phoneme_mask_prob=float(dataset_params.get("phoneme_mask_prob", 0.8)),
                       replace_prob=float(dataset_params.get("replace_prob", 0.1)))
        self.epoch = 0

    def __len__(self) -> int:
        return self.num_batches

    def __iter__(self):
        from .batch import synthetic_collate_batch
        base = self.seed + (100003 * self.epoch if self.redraw else 0)   # validation batches stay fixed
        self.epoch += 1
        for i in range(self.num_batches):
            yield synthetic_collate_batch(self.batch_size, seed=base + i, **self.kw)


def setup_dataset_and_dataloader(config, device, use_token_ids: bool = False, num_tokens: int = 64000, num_workers: int = 4):
    """R/train.py:240-257.  Returns ``(train_loader, val_loader)`` yielding GLOBAL collate batches.

    ``training_dataset``: ``synthetic`` = the built-in tensor generator (no per-sample Python at all); ``synthetic_records`` =
    synthetic dataset rows through the reference's masking + collate pipeline in ``num_workers`` background processes;
    anything else = a ``datasets.load_dataset`` name whose rows go through that same pipeline (``plbert_b200.data``: mirrors
    of R/dataloader.py's classes plus the worker pool — the reference runs them with ``num_workers=0``, R/train.py:253)."""
    tp = config["training_params"]
    name = str(tp.get("training_dataset", "")).lower()
    if name == "synthetic":
        n_train = int(tp.get("synthetic_batches_per_epoch", 64))
        n_val = int(tp.get("synthetic_val_batches", 2))
        mk = lambda n, seed, redraw: SyntheticLoader(int(tp["batch_size"]), config["dataset_params"], n, seed, use_token_ids,
                                                     num_tokens, redraw=redraw)
        return mk(n_train, 1234, True), mk(n_val, 987654, False)
    from .data import build_dataloader, synthetic_records
    if name == "synthetic_records":
        dataset = synthetic_records(int(tp.get("synthetic_rows", 4096)), seed=1234, num_tokens=num_tokens)
    else:
        try:
            from datasets import load_dataset
        except ImportError as e:
            raise RuntimeError("training_dataset names a datasets.load_dataset source but `datasets` is not importable: " + str(e)) from e
        dataset = load_dataset(tp["training_dataset"], split=tp["split"])
    return build_dataloader(dataset, batch_size=tp["batch_size"], device=device, dataset_config=config["dataset_params"],
                            use_token_ids=use_token_ids, num_workers=num_workers)


def _num_symbols() -> int:
    try:
        from char_indexer import symbols  # R/char_indexer.py:15 when the reference tree is on PYTHONPATH
        return len(symbols)
    except ImportError:
        from .batch import NUM_SYMBOLS
        return NUM_SYMBOLS


def initialize_model(config, log_dir: str, resuming: bool, device, use_token_ids: bool = False,
                     num_tokens: Optional[int] = None):
    """R/train.py:259-286: model + optimizer, optional ``pretrained_model`` warm start, resume from the newest
    ``step_*`` file of the run directory.  Returns ``(model, optimizer, current_step)``."""
    from transformers import AlbertConfig, AlbertModel
    from .model import MultiTaskModel, PhonemeOnlyModel
    from .optim import FusedAdamW
    mp = dict(config["model_params"])
    pretrained = mp.pop("pretrained_model", None)
    mp.pop("dropout", None)  # not an AlbertConfig field: inert in the reference
    n_sym = _num_symbols()
    albert = AlbertModel(AlbertConfig(vocab_size=n_sym, **mp))
    if use_token_ids:
        if not num_tokens:
            raise ValueError("--use_token_ids needs --num_tokens (the tokenizer's vocabulary size)")
        model = MultiTaskModel(albert, num_phonemes=n_sym, num_tokens=num_tokens, hidden_size=mp["hidden_size"])
    else:
        model = PhonemeOnlyModel(albert, num_phonemes=n_sym, hidden_size=mp["hidden_size"])
    model = model.to(device)
    optimizer = FusedAdamW(model.parameters(), lr=float(config["training_params"]["learning_rate"]))
    model._ensure_engine(torch.device(device))  # arenas exist before any optimizer state is loaded into them
    if pretrained:
        print(f"Loading pretrained model from: {pretrained}")
        load_checkpoint(model, optimizer, pretrained, device=device)
    found, current_step = find_latest_checkpoint(log_dir)
    if found and resuming:
        path = next(os.path.join(log_dir, f) for f in sorted(os.listdir(log_dir)) if f.startswith(f"step_{current_step}."))
        load_checkpoint(model, optimizer, path, device=device)
    else:
        current_step = 0
    return model, optimizer, current_step


@torch.no_grad()
def validate(model, val_dataloader: Iterable, rank: int = 0, world: int = 1) -> float:
    """R/train.py:288-305: mean batch loss over the validation loader (no gradients, weights untouched); with several
    ranks every rank scores its slice of each batch and the sample-weighted mean is all-reduced."""
    total, n = 0.0, 0
    for batch in val_dataloader:
        if world == 1:
            total += float(model.loss(batch).item())
        else:
            mine = shard_batch(batch, rank, world, strict=False)
            if mine is not None:   # uneven last batch: a share may be empty; shares are weighted by their sample count
                total += float(model.loss(mine).item()) * len(mine[-2]) / len(batch[-2])
        n += 1
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([total], dtype=torch.float64, device=next(model.parameters()).device)
        dist.all_reduce(t)
        total = float(t.item())
    return total / max(n, 1)


class MetricsLogger:
    """wandb.log when available (R/train.py:228-229,392-410), else one JSON object per line in ``metrics.jsonl``."""

    def __init__(self, log_dir: str, config, enabled: bool, use_wandb: bool):
        self.enabled, self.wandb, self.file = enabled, None, None
        if not enabled:
            return
        if use_wandb:
            try:
                import wandb
                wandb.init(project="pl_bert_phoneme_only_diacritized", config=config)
                self.wandb = wandb
            except Exception as e:  # not installed / offline: fall through to the file
                print(f"wandb unavailable ({type(e).__name__}); logging to {log_dir}/metrics.jsonl")
        if self.wandb is None:
            self.file = open(os.path.join(log_dir, "metrics.jsonl"), "a")

    def log(self, row: Dict[str, object]) -> None:
        if not self.enabled:
            return
        if self.wandb is not None:
            self.wandb.log(row)
        else:
            self.file.write(json.dumps(row) + "\n")
            self.file.flush()

    def close(self) -> None:
        if self.file is not None:
            self.file.close()


def train_loop(model, optimizer, train_dataloader, val_dataloader, current_step: int, num_steps: int, save_interval: int,
               log_interval: int, log_dir: str, logger: MetricsLogger, max_epochs: int = 10, rank: int = 0, world: int = 1,
               checkpoint_extension: str = "t7", log_lag: int = 0) -> Tuple[int, int]:
    """R/train.py:338-379 with the fused step: validation first, then per batch ``BatchFeeder.step`` (forward + loss +
    backward on the batch staged during the previous step) and ``optimizer.step``; checkpoint + validation every
    ``save_interval`` steps; stops at ``num_steps`` or after ``max_epochs`` passes.

    ``log_lag`` = 0 reads every step's loss back right away like the reference (``loss.item()``, R/train.py:395 — one
    device->host sync per step); ``log_lag`` = k logs a step's loss k steps later (the value is kept in a device scalar
    meanwhile), so the host never waits for the step it has just enqueued.  Every loss is still logged, in order; the
    queue is drained before checkpoints, validation and the end of training."""
    from .batch import BatchFeeder
    main = rank == 0
    say = print if main else (lambda *a, **k: None)
    losses: deque = deque(maxlen=log_interval)

    def run_validation(epoch: int) -> float:
        val = validate(model, val_dataloader, rank, world)
        logger.log({"val_phoneme_loss": val, "step": current_step, "epoch": epoch})
        say(f"Validation at step {current_step}: Phoneme Loss: {val:.4f}")
        return val

    pending: deque = deque()   # (epoch, device loss) of steps whose loss has not been read back yet

    def flush(keep: int) -> None:
        while len(pending) > keep:
            epoch, t = pending.popleft()
            value = float(t.item())
            losses.append(value)
            row = {"phoneme_loss": value, "epoch": epoch}
            if len(losses) == log_interval:
                row["phoneme_loss_avg"] = sum(losses) / len(losses)
            logger.log(row)

    current_epoch = 0
    run_validation(current_epoch)
    while current_epoch < max_epochs:
        current_epoch += 1
        say(f"Starting epoch {current_epoch}")
        feed = BatchFeeder(model, (shard_batch(b, rank, world) for b in train_dataloader))
        while not feed.exhausted:
            loss = feed.step()          # process_batch + backward (R/train.py:351-356); next batch staged underneath
            optimizer.step()
            current_step += 1
            # the engine's loss scalar is overwritten by the next step: a lagged read needs its own copy
            pending.append((current_epoch, loss.clone() if log_lag > 0 else loss))
            flush(log_lag)
            if current_step % save_interval == 0:
                flush(0)
                if main:
                    say("Saving..")
                    path = save_checkpoint(model, optimizer, current_step, log_dir, current_epoch, extension=checkpoint_extension)
                    say(f"Checkpoint saved at: {path}")
                run_validation(current_epoch)
            if current_step >= num_steps:
                flush(0)
                return current_step, current_epoch
        flush(0)
    return current_step, current_epoch


def train(args: Optional[Dict[str, object]] = None) -> Tuple[int, int]:
    """R/train.py:133-162."""
    if args is None:
        args = parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise _abi.PlbError("plbert_b200.train needs a CUDA device (B200); there is no CPU training path")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    if rank == 0:
        decided = [setup_config_and_directories(args, str(args["config_path"]))]
    else:
        decided = [None]
    if world > 1:
        # rank 0 alone decides fresh-vs-resume (it may purge or create the run directory, and once it has copied the config
        # in, every later look at the directory says "resume"); the other ranks take its decision as is
        import torch.distributed as dist
        dist.broadcast_object_list(decided, src=0)
    config, log_dir, resuming = decided[0]
    tp = config["training_params"]
    print(f"Resuming training from '{log_dir}' with existing config." if resuming else f"Starting new training run in '{log_dir}'.")
    logger = MetricsLogger(log_dir, config, enabled=rank == 0, use_wandb=not args.get("no_wandb", False))
    # optional block of the YAML that only this implementation reads (the reference's keys stay untouched):
    #   b200_params: {use_token_ids, num_tokens, token_rows: masked|all_valid, pack_sequences, checkpoint_extension, log_lag,
    #                 num_workers (input pipeline processes)}
    bp = dict(config.get("b200_params") or {})
    use_tok = bool(args.get("use_token_ids") or bp.get("use_token_ids", False))
    num_tokens = args.get("num_tokens") or bp.get("num_tokens")
    train_dl, val_dl = setup_dataset_and_dataloader(config, device, use_tok, int(num_tokens or 64000),
                                                    num_workers=int(bp.get("num_workers", 4)))
    model, optimizer, current_step = initialize_model(config, log_dir, resuming, device, use_tok, num_tokens)
    if "pack_sequences" in bp:
        model.pack_sequences = bool(bp["pack_sequences"])
    if use_tok and "token_rows" in bp:
        if bp["token_rows"] not in ("masked", "all_valid"):
            raise ValueError("b200_params.token_rows must be 'masked' or 'all_valid'")
        model.token_rows = bp["token_rows"]
    if world > 1:
        from .ddp import GradientSynchronizer
        sync = GradientSynchronizer(model._engine)   # broadcasts rank 0's weights, hooks the bucketed all-reduce
        optimizer.grad_scale = sync.grad_scale
    if rank == 0:
        print("Start training...")
    t0 = time.time()
    step, epoch = train_loop(model, optimizer, train_dl, val_dl, current_step, int(tp["num_steps"]), int(tp["save_interval"]),
                             int(tp["log_interval"]), log_dir, logger, rank=rank, world=world,
                             checkpoint_extension=str(bp.get("checkpoint_extension") or args.get("checkpoint_extension", "t7")),
                             log_lag=int(bp.get("log_lag", 0)))
    torch.cuda.synchronize()
    if rank == 0:
        print(f"Training completed at step {step}, epoch {epoch} ({time.time() - t0:.1f} s)")
    logger.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return step, epoch


if __name__ == "__main__":
    train()
