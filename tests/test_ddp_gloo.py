"""Data-parallel plumbing on CPU: two gloo ranks (world_size 2) exercise the same bucket code the NCCL path uses
(plbert_b200/ddp.py) and the reference's DDP semantics (per-rank loss normalisation + gradient MEAN)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import plbert_oracle as O
from plbert_b200 import ddp
from plbert_b200.batch import synthetic_collate_batch
from plbert_b200.engine import EngineConfig, ParamLayout

CFG = O.OracleConfig(vocab_size=188, embedding_size=16, hidden_size=32, num_attention_heads=2, intermediate_size=64,
                     num_hidden_layers=2, max_position_embeddings=64)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _flat(layout, grads):
    flat = torch.zeros(layout.total, dtype=torch.float64)
    for n, g in grads.items():
        o = layout.offsets[n]
        flat[o:o + g.size] = torch.from_numpy(np.ascontiguousarray(g).reshape(-1))
    return flat


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        layout = ParamLayout(EngineConfig(188, 16, 32, 2, 64, 2, 64, 2, 1e-12, 0, 188, 50))
        # parameter broadcast: every rank ends with rank 0's weights
        params = O.init_params(CFG, 188, 50, seed=100 + rank)
        flat_p = _flat(layout, params)
        ddp.broadcast_params(flat_p, 0)
        for n in params:
            o = layout.offsets[n]
            params[n] = flat_p[o:o + params[n].size].numpy().reshape(params[n].shape).copy()
        # each rank: its own half of the global batch, its own loss normalisation (R/train.py:128-129)
        tok, lab, msk, lens, midx = synthetic_collate_batch(4, 48, seed=9, with_token_ids=True, num_tokens=50, word_pred_prob=0.4)
        sl = slice(rank * 2, rank * 2 + 2)
        loss, _, grads, _ = O.loss_and_grads(params, CFG, msk[sl].numpy(), lab[sl].numpy(), lens[sl], midx[sl], token_labels=tok[sl].numpy())
        flat_g = _flat(layout, grads)
        bounds = ddp.bucket_bounds(layout.head_end, layout.trainable_end, max_bucket_elems=4096)
        assert bounds[0][0] == 0 and bounds[-1][1] == layout.trainable_end
        assert all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
        works = ddp.all_reduce_buckets(flat_g, bounds, async_op=True)
        for w in works:
            w.wait()
        flat_g /= world  # folded into AdamW's grad_scale on the CUDA path
        if rank == 0:
            full_loss, _, full_grads, _ = O.loss_and_grads(params, CFG, msk.numpy(), lab.numpy(), lens, midx, token_labels=tok.numpy())
            ref = _flat(layout, full_grads)
            assert all(len(m) > 0 for m in midx), "equivalence needs every sample to carry masked rows (SURVEY.md §8e)"
            err = ((flat_g - ref).norm() / ref.norm()).item()
            out.put(("ok", err, float(loss), float(full_loss)))
        dist.barrier()
    except Exception as e:  # pragma: no cover
        if rank == 0:
            out.put(("error", repr(e), 0, 0))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gradient_mean_equals_single_rank_on_concatenated_batch():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    status, err, loss0, full = out.get(timeout=240)
    for p in procs:
        p.join(60)
    assert status == "ok", err
    assert all(p.exitcode == 0 for p in procs)
    assert err < 1e-12, err  # float64: the two formulations are the same sum


def test_bucket_bounds():
    assert ddp.bucket_bounds(0, 100) == [(0, 100)]
    assert ddp.bucket_bounds(64, 200) == [(0, 64), (64, 200)]
    assert ddp.bucket_bounds(64, 200, 50) == [(0, 50), (50, 64), (64, 114), (114, 164), (164, 200)]


def test_head_bounds_without_fused_weights():
    """Opt-in fused WGRAD + all-reduce (ddp.GradientSynchronizer(fuse_head_allreduce=True)): the NCCL head bucket must cover
    exactly the head region minus the fused weights, in order, without gaps inside the remaining parameters."""
    layout = ParamLayout(EngineConfig(188, 16, 32, 2, 64, 2, 64, 2, 1e-12, 0, 188, 50))
    names = ["token_predictor.weight"]
    b = ddp.head_bounds_without(layout, names)
    o, n = layout.offsets["token_predictor.weight"], layout.numel("token_predictor.weight")
    assert o == 0 and b == [(n, layout.head_end)]
    covered = torch.zeros(layout.head_end, dtype=torch.int32)
    for lo, hi in b:
        covered[lo:hi] += 1
    covered[o:o + n] += 1
    assert (covered == 1).all()
    for name in ("token_predictor.bias", "phoneme_predictor.weight", "phoneme_predictor.bias"):
        lo = layout.offsets[name]
        assert any(a <= lo and lo + layout.numel(name) <= e for a, e in b), name
    assert ddp.head_bounds_without(layout, []) == [(0, layout.head_end)]
    both = ddp.head_bounds_without(layout, ["phoneme_predictor.weight", "token_predictor.weight"])
    assert all(a < e for a, e in both) and both[-1][1] == layout.head_end
    assert sum(e - a for a, e in both) == layout.head_end - n - layout.numel("phoneme_predictor.weight")


def _sync_worker(rank, world, port, out):
    """The real ``GradientSynchronizer`` hooks + ``Engine._empty_step`` on CPU tensors: rank 1's shard has no masked rows."""
    import types
    from plbert_b200.engine import Engine
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        layout = ParamLayout(EngineConfig(188, 16, 32, 2, 64, 2, 64, 2, 1e-12, 0, 188, 50))
        eng = types.SimpleNamespace(layout=layout, device=torch.device("cpu"), shadow_valid=True, always_reduce=False,
                                    have_grads=False, loss_buf=torch.ones(1),
                                    params=torch.full((layout.total,), float(rank)), grads=torch.zeros(layout.total),
                                    _before_head_backward=lambda: None, _after_head_backward=lambda: None,
                                    _after_encoder_backward=lambda: None)
        sync = ddp.GradientSynchronizer(eng)
        assert eng.always_reduce and not eng.shadow_valid and sync.grad_scale == 0.5
        assert not eng.params.any()                                    # rank 0's weights everywhere
        assert sync.symm is None and not sync.fused_heads             # CPU tensors: the NCCL / gloo bucket path
        if rank == 0:   # a normal step: gradients, then the hooks in backward order
            eng.grads[:layout.trainable_end] = 2.0
            eng.have_grads = True
            eng._before_head_backward()
            eng._after_head_backward()
            eng._after_encoder_backward()
        else:           # empty shard (R/train.py:129): must still enter both collectives with zeros, and step
            Engine._empty_step(eng)
        g = eng.grads * sync.grad_scale                                # AdamW folds 1/world in
        ok = eng.have_grads and bool((g[:layout.trainable_end] == 1.0).all()) and not g[layout.trainable_end:].any()
        if rank == 1:
            ok = ok and eng.loss_buf.item() == 0.0
        res = torch.tensor([1.0 if ok else 0.0])
        dist.all_reduce(res)
        if rank == 0:
            out.put(("ok" if res.item() == world else "mismatch", res.item()))
    except Exception as e:  # pragma: no cover
        if rank == 0:
            out.put(("error", repr(e)))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_empty_rank_joins_the_collectives():
    """A rank without masked rows must not return before the bucket all-reduces (the peers would hang) and must run the
    same optimizer step on the averaged gradient."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    status, detail = out.get(timeout=240)
    for p in procs:
        p.join(60)
    assert status == "ok", detail
    assert all(p.exitcode == 0 for p in procs)


def test_single_process_empty_step_skips_the_optimizer():
    import types
    from plbert_b200.engine import Engine
    eng = types.SimpleNamespace(always_reduce=False, have_grads=True, loss_buf=torch.ones(1))
    Engine._empty_step(eng)
    assert eng.loss_buf.item() == 0.0 and not eng.have_grads
