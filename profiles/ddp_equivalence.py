#!/usr/bin/env python
"""N-rank data-parallel step through the REAL NCCL path (ddp.GradientSynchronizer: side stream, two buckets, 1/world folded
into FusedAdamW) == one-rank step on the concatenated batch (R/train.py:160-162,218-221 DDP semantics).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        profiles/ddp_equivalence.py [--fuse] > profiles/r2_ddp_equivalence.json

--fuse: GradientSynchronizer(fuse_head_allreduce=True) — the token head's WGRAD adds into every rank's arena through the
NVSwitch multicast mapping (plb_gemm_wgrad_allreduce) instead of the NCCL head bucket.

Every rank takes its contiguous share of each global batch (train.shard_batch), runs 3 optimizer steps; rank 0 then repeats
the 3 steps alone on the whole batches from the same initial weights and compares
  * the averaged gradient of step 1 (cosine per tensor, worst reported),
  * the weights after 3 steps (cosine of the UPDATE w3 - w0 per tensor; lr 1e-3 so that the weights really move; bar 0.99:
    Adam normalises every element's step to ~lr, so elements whose gradient sits at the fp32 summation-order noise floor
    may step the other way and the two trajectories then drift apart — the step-1 gradients are the exact comparison),
  * that all ranks hold bit-identical weights at the end.
Case "empty_rank": the last rank's share has no masked rows at all (R/train.py:129) — it must still join both all-reduces and
step (no hang, identical replicas); the averaged gradient is then (1/world) x the gradient of the other ranks' mean.
Equivalence with the concatenated batch needs every rank to see the same number of non-empty samples (the reference's
per-rank loss normalisation), which the "equal" case guarantees by construction."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DIMS = dict(hidden_size=768, num_attention_heads=12, intermediate_size=2048, max_position_embeddings=512, num_hidden_layers=12)
VP, VT, PER_RANK, SEQ, STEPS, LR = 188, 64000, 4, 512, 3, 1e-3


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-300))


def build(dev, state=None):
    import plbert_b200 as P
    from transformers import AlbertConfig, AlbertModel
    torch.manual_seed(0)
    model = P.MultiTaskModel(AlbertModel(AlbertConfig(vocab_size=VP, **DIMS)), VP, VT, 768)
    if state is not None:
        model.load_state_dict(state)
    model = model.to(dev)
    return model, P.FusedAdamW(model.parameters(), lr=LR)


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from plbert_b200.batch import synthetic_collate_batch
    from plbert_b200.ddp import GradientSynchronizer
    from plbert_b200.train import shard_batch
    report = {"world": world, "dims": DIMS, "num_tokens": VT, "per_rank_batch": PER_RANK, "seq": SEQ, "steps": STEPS, "lr": LR,
              "cases": {}}
    for case in ("equal", "empty_rank"):
        batches = []
        for i in range(STEPS):
            tok, lab, msk, lens, midx = synthetic_collate_batch(PER_RANK * world, SEQ, seed=700 + i, with_token_ids=True, num_tokens=VT)
            assert all(len(m) > 0 for m in midx)
            if case == "empty_rank":
                midx = [m if b < PER_RANK * (world - 1) else [] for b, m in enumerate(midx)]
            batches.append((tok, lab, msk, lens, midx))
        model, opt = build(dev)
        init = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        eng = model._ensure_engine(dev)
        sync = GradientSynchronizer(eng, fuse_head_allreduce="--fuse" in sys.argv)
        opt.grad_scale = sync.grad_scale
        report["fused_head_allreduce"] = {"heads": sync.fused_heads, "why_not": sync.fuse_error}
        grads1, losses = None, []
        for i, gb in enumerate(batches):
            loss = model.loss_and_backward(shard_batch(gb, rank, world))
            if i == 0:
                torch.cuda.synchronize()
                grads1 = {n: eng.g(n).detach().clone() * sync.grad_scale for n in eng.layout.offsets if n not in eng.layout.frozen}
            opt.step()
            losses.append(float(loss.item()))
        torch.cuda.synchronize()
        # replicas must stay bit-identical
        flat = eng.params[:eng.layout.trainable_end].clone()
        ref = flat.clone()
        dist.broadcast(ref, 0)
        same = torch.tensor([1.0 if torch.equal(flat, ref) else 0.0], device=dev)
        dist.all_reduce(same)
        if rank == 0:
            one, opt1 = build(dev, init)
            e1 = one._ensure_engine(dev)
            g_ref, l_ref = None, []
            for i, gb in enumerate(batches):
                if case == "empty_rank":   # what the ranks with data computed; the averaged gradient is (world-1)/world of it
                    gb = tuple(t[:PER_RANK * (world - 1)] for t in gb)
                l = one.loss_and_backward(gb)
                if i == 0:
                    torch.cuda.synchronize()
                    g_ref = {n: e1.g(n).detach().clone() for n in e1.layout.offsets if n not in e1.layout.frozen}
                opt1.step()
                l_ref.append(float(l.item()))
            scale = (world - 1) / world if case == "empty_rank" else 1.0
            worst_g, worst_u, worst_ratio = 1.0, 1.0, 0.0
            for n in grads1:
                if n.endswith("attention.key.bias"):
                    continue
                worst_g = min(worst_g, cosine(grads1[n], g_ref[n]))
                worst_ratio = max(worst_ratio, abs(float(grads1[n].norm() / (scale * g_ref[n].norm())) - 1.0))
                if case == "equal":
                    w0 = init[n].to(dev)
                    worst_u = min(worst_u, cosine(eng.p(n) - w0, e1.p(n) - w0))
            rec = {"worst_gradient_cosine_step1": worst_g, "worst_gradient_norm_ratio_error": worst_ratio,
                   "replicas_bit_identical": bool(same.item() == world), "rank0_losses": losses, "one_rank_losses": l_ref}
            if case == "equal":
                rec["worst_update_cosine_after_3_steps"] = worst_u
            rec["pass"] = bool(worst_g >= 0.9999 and worst_ratio < 1e-2 and rec["replicas_bit_identical"]
                               and (case != "equal" or worst_u >= 0.99))
            report["cases"][case] = rec
            del one, opt1, e1
        del model, opt, eng, sync
        torch.cuda.empty_cache()
        dist.barrier()
    if rank == 0:
        report["pass"] = all(c["pass"] for c in report["cases"].values())
        print(json.dumps(report, indent=1), flush=True)
    dist.destroy_process_group()
    if rank == 0 and not report["pass"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
