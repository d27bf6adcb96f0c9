#!/usr/bin/env python
"""Benchmark of the PL-BERT training hot path on B200 (contract: see the task brief / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--model multitask|phoneme_only]

One "step" = forward + masked-row loss (both heads) + backward + gradient all-reduce (N > 1) + AdamW over one batch
of BASELINE.json configs[1]: ALBERT-base from R/configs/config.yml (hidden 768, 12 shared layers, 12 heads, FFN 2048,
vocab 188), bf16, 192 x 512 synthetic phoneme tokens per GPU with the reference's whole-word 15 % masking.
``value`` counts valid phoneme tokens per second over all ranks with the batch already resident in HBM; ``e2e`` is
the same step through the public API (``BatchFeeder.step`` + ``FusedAdamW.step`` + ``loss.item()``) from host buffers:
every step flattens its collate tuple, copies it pinned -> device and reads the loss back.
``--impl reference`` times the CPU port of the reference step (oracle/reference_port.py) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL_PARAMS = dict(hidden_size=768, num_attention_heads=12, intermediate_size=2048, max_position_embeddings=512,
                    num_hidden_layers=12)  # R/configs/config.yml:32-39
NUM_PHONEMES = 188   # len(symbols), R/char_indexer.py:15
NUM_TOKENS = 64000   # aubmindlab/bert-base-arabertv2 vocab (R/convert_to_hf.py:24,32) — assumption, no network here
BATCH, SEQ = 192, 512
CPU_SAMPLE_BATCH = 8  # BASELINE.json configs[0]: the reference's own CPU-runnable case (8 x 512)
METRIC = "phoneme_tokens_per_sec_training"


def step_flops(lengths, n_rows, two_heads):
    """Algorithmic FLOPs of one step: GEMMs 2MNK, backward = 2x forward, no recompute counted."""
    E, H, I, NL = 128, MODEL_PARAMS["hidden_size"], MODEL_PARAMS["intermediate_size"], MODEL_PARAMS["num_hidden_layers"]
    T = sum(lengths)
    sq = sum(l * l for l in lengths)
    f = 3 * (T * (2 * E * H + NL * (8 * H * H + 4 * H * I)) + NL * 4 * H * sq)
    f += 3 * 2 * H * NUM_PHONEMES * n_rows
    if two_heads:
        f += 3 * 2 * H * NUM_TOKENS * n_rows
    return f


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(burst=d["bf16_tflops"], sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), hbm=d["hbm_gbs"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def gemm_dram_traffic(two_heads):
    """Average DRAM bytes (read + write) per GEMM launch of one step, from the newest committed ncu capture of the same
    step (profiles/r1*_gemm_dram_traffic*.json; `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`, reduced by
    profiles/gemm_dram_traffic.py).  None if absent."""
    import glob
    suffix = "_multitask.json" if two_heads else ".json"
    files = sorted(f for f in glob.glob(os.path.join(ROOT, "profiles", "r1*_gemm_dram_traffic*.json"))
                   if f.endswith("_gemm_dram_traffic" + suffix))
    for path in reversed(files):
        try:
            with open(path) as f:
                return json.load(f)["avg_dram_bytes_per_launch"]
        except (OSError, KeyError, ValueError):
            continue
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        power = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "", 1).isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "power_w_max": max(power) if power else None}


def make_batch(seed, batch, two_heads):
    from plbert_b200.batch import synthetic_collate_batch
    return synthetic_collate_batch(batch, SEQ, seed=seed, with_token_ids=two_heads, num_tokens=NUM_TOKENS)


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: the reference step (port) on the host cores, bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import reference_port as RP
    two_heads = args.model == "multitask"
    sample_batch = CPU_SAMPLE_BATCH
    # bound the run: probe one step, shrink the per-step sample if K + W steps would not fit in ~4 minutes
    probe, _, _ = RP.time_steps(make_batch(1234, sample_batch, two_heads), NUM_TOKENS if two_heads else None, 1, 0)
    while sample_batch > 1 and probe * (args.steps + args.warmup) * sample_batch / CPU_SAMPLE_BATCH > 240:
        sample_batch //= 2
    batch = make_batch(1234, sample_batch, two_heads)
    dt, tokens, threads = RP.time_steps(batch, NUM_TOKENS if two_heads else None, args.steps, args.warmup)
    value = tokens / dt
    sample = f"{sample_batch}x{SEQ} tokens per step (BASELINE.json configs[0] shape), fp32, {args.steps} timed steps"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic",
            "config": {"workload": workload_name(two_heads), "sample": sample, "device": "cpu"},
            "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_name(two_heads):
    heads = f"phoneme head V={NUM_PHONEMES} + token head V={NUM_TOKENS} (MultiTaskModel)" if two_heads else \
        f"phoneme head V={NUM_PHONEMES} (PhonemeOnlyModel)"
    return (f"BASELINE.json configs[1]: PL-BERT config.yml ALBERT (hidden 768, 12 shared layers, 12 heads, ffn 2048), {heads}, "
            f"one bf16 train step (fwd + masked CE + bwd + AdamW), {BATCH}x{SEQ} synthetic phoneme tokens per GPU, 15% word masking")


def run_b200(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import plbert_b200 as P
    from plbert_b200.ddp import GradientSynchronizer
    from transformers import AlbertConfig, AlbertModel

    two_heads = args.model == "multitask"
    torch.manual_seed(0)
    enc = AlbertModel(AlbertConfig(vocab_size=NUM_PHONEMES, **MODEL_PARAMS))
    model = (P.MultiTaskModel(enc, NUM_PHONEMES, NUM_TOKENS, 768) if two_heads else P.PhonemeOnlyModel(enc, NUM_PHONEMES, 768)).to(dev)
    opt = P.FusedAdamW(model.parameters(), lr=7e-5)
    heads = ["phoneme_predictor", "token_predictor"] if two_heads else ["phoneme_predictor"]

    n_batches = 2
    host_batches = [make_batch(1234 + rank + 1000 * i, BATCH, two_heads) for i in range(n_batches)]
    eng = model._ensure_engine(dev)
    sync = GradientSynchronizer(eng) if world > 1 else None
    if sync is not None:
        opt.grad_scale = sync.grad_scale
    lr, gs = 7e-5, opt.grad_scale

    def stage(hb):
        if two_heads:
            tok, lab, msk, lens, midx = hb
            return model._stager.stage(msk, lens, midx, [lab, tok])
        lab, msk, lens, midx = hb
        return model._stager.stage(msk, lens, midx, [lab])

    # resident batches for the device-only number (their own stager so the slots are not recycled)
    from plbert_b200.batch import BatchStager
    resident = []
    for hb in host_batches:
        st = BatchStager(dev)
        model._stager = st
        resident.append(stage(hb))
    model._stager = BatchStager(dev)
    tokens_per_step = [b.num_valid_tokens for b in resident]
    rows_per_step = [b.M for b in resident]

    def device_step(i):
        b = resident[i % n_batches]
        eng.loss_and_backward(b, heads)
        eng.adamw_step(lr, grad_scale=gs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        device_step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    l0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        device_step(i)
    e1.record()
    barrier()
    launches = eng.launches - l0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    tok = torch.tensor([float(sum(tokens_per_step[i % n_batches] for i in range(args.steps)))], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tok, op=dist.ReduceOp.SUM)
    ms_total, total_tokens = t.item(), tok.item()
    value = total_tokens / (ms_total * 1e-3)

    # ---- end to end through the public API: host batch -> pinned -> H2D -> step -> loss D2H ---------------
    # BatchFeeder = the input pipeline of the public API: step i runs on the batch staged during step i-1 and stages
    # batch i+1 (host flatten -> pinned buffer -> H2D) while the GPU works; every step still uploads its own batch.
    import itertools
    from plbert_b200.batch import BatchFeeder
    feed = BatchFeeder(model, (host_batches[i % n_batches] for i in itertools.count()))

    def api_step(i):
        loss = feed.step()
        opt.step()
        return loss.item()

    for i in range(2):
        api_step(i)
    barrier()
    e0.record()
    for i in range(args.steps):
        last_loss = api_step(i)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total_tokens / (t.item() * 1e-3)
    h2d = model.last_batch.h2d_bytes

    # ---- roofline of the dominant kernel (the tcgen05 GEMM): CUDA events around every GEMM launch of one more step ----
    roof = None
    records = []
    orig = eng._gemm

    def timed(op, epi, M, N, K, *a, **k):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        orig(op, epi, M, N, K, *a, **k)
        e.record()
        records.append((2.0 * M * N * K, s, e))

    eng._gemm = timed
    device_step(0)  # every rank runs it (the step contains collectives); rank 0 reports
    torch.cuda.synchronize()
    eng._gemm = orig
    if rank == 0:
        fl = sum(r[0] for r in records)
        tt = sum(r[1].elapsed_time(r[2]) for r in records) * 1e-3
        pk = measured_peaks()
        achieved = fl / tt / 1e12
        roof = {"bound": "tensor", "kernel": "gemm_kernel (tcgen05/TMA; fwd+dgrad+wgrad, all epilogues)", "achieved": achieved,
                "peak": pk["sustained"], "unit": "TFLOP/s", "frac": achieved / pk["sustained"],
                "traffic": gemm_dram_traffic(two_heads),
                "launches_per_step": len(records), "avg_launch_us": tt / len(records) * 1e6,
                "flops_per_launch": fl / len(records), "gemm_share_of_step": tt / (ms_total * 1e-3 / args.steps),
                "peak_source": pk["source"] + ", sustained figure (kernel timed inside a long step); burst " + str(pk["burst"])}
    if world > 1:
        barrier()

    if rank == 0:
        pk = measured_peaks()
        fl_step = sum(step_flops([SEQ] * BATCH, rows_per_step[i % n_batches], two_heads) for i in range(args.steps)) / args.steps
        step_s = ms_total * 1e-3 / args.steps
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import reference_port as RP
            cb = make_batch(1234, CPU_SAMPLE_BATCH, two_heads)
            dt, ctok, threads = RP.time_steps(cb, NUM_TOKENS if two_heads else None, 2, 1)
            cpu = {"value": ctok / dt, "unit": "tokens/s", "cores": threads, "kind": "port",
                   "sample": f"{CPU_SAMPLE_BATCH}x{SEQ} tokens per step, fp32, 1 warm-up + 2 timed steps of the reference step "
                             f"(oracle/reference_port.py: HF AlbertModel + heads + loss loop + torch AdamW)"}
        line = {"metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": workload_name(two_heads), "parallelism": f"dp{world}",
                           "l2": "inputs larger than L2: each step streams ~24 GB of bf16 activations (126 MB L2), two batches alternate",
                           "tokens_per_step_per_gpu": tokens_per_step[0], "masked_rows_per_step": rows_per_step[0],
                           "loss": last_loss},
                "per_gpu": value / world,
                "tensor_pipe_frac_of_bf16_peak": {"algorithmic_tflop_per_step": fl_step / 1e12,
                                                  "achieved_tflops": fl_step / step_s / 1e12,
                                                  "vs_burst": fl_step / step_s / 1e12 / pk["burst"],
                                                  "vs_sustained": fl_step / step_s / 1e12 / pk["sustained"]},
                "roofline": roof, "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
                "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="multitask", choices=["multitask", "phoneme_only"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_b200(args)


if __name__ == "__main__":
    main()
