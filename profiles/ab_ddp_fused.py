"""N-rank A/B of the head-gradient all-reduce: fused into the WGRAD epilogue over NVSwitch multicast (plb_gemm_wgrad_allreduce)
vs the NCCL bucket on a side stream.  Two engines in one process per rank, blocks of steps in alternation, max over ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 profiles/ab_ddp_fused.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import plbert_b200 as P
    from plbert_b200.ddp import GradientSynchronizer
    from transformers import AlbertConfig, AlbertModel
    heads = ["phoneme_predictor", "token_predictor"]
    arms = {}
    for fused in (False, True):
        torch.manual_seed(0)
        enc = AlbertModel(AlbertConfig(vocab_size=bench.NUM_PHONEMES, **bench.MODEL_PARAMS))
        model = P.MultiTaskModel(enc, bench.NUM_PHONEMES, bench.NUM_TOKENS, 768).to(dev)
        opt = P.FusedAdamW(model.parameters(), lr=7e-5)
        eng = model._ensure_engine(dev)
        sync = GradientSynchronizer(eng, fuse_head_allreduce=fused)
        batches = [model.stage_batch(bench.make_batch(1234 + 7 * rank + i, bench.BATCH, True), stager=model.new_stager()) for i in range(2)]
        arms[fused] = (model, eng, sync, batches)
        if rank == 0:
            print(f"arm fused={fused}: fused heads {sync.fused_heads} ({sync.fuse_error})", flush=True)

    def block(fused, n=8):
        model, eng, sync, batches = arms[fused]
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            eng.loss_and_backward(batches[i % 2], heads)
            eng.adamw_step(7e-5, grad_scale=sync.grad_scale)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), float(eng.loss_buf[0].item())

    for fused in (False, True):
        block(fused, 3)
    res = {False: [], True: []}
    for rnd in range(4):
        for fused in (False, True):
            ms, loss = block(fused)
            res[fused].append(ms)
            if rank == 0:
                print(f"round {rnd} fused={fused}: {ms:.3f} ms/step (max over ranks)  loss {loss:.4f}", flush=True)
    if rank == 0:
        a, b = sum(res[False]) / 4, sum(res[True]) / 4
        print(json.dumps({"world": world, "nccl_bucket_ms": a, "fused_ms": b, "gain_pct": (a - b) / a * 100}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
