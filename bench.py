#!/usr/bin/env python
"""Benchmark of the PL-BERT training hot path on B200 (contract: see the task brief / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference|library] [--model multitask|phoneme_only]

One "step" = forward + masked-row loss (both heads) + backward + gradient all-reduce (N > 1) + AdamW over one batch
of BASELINE.json configs[1]: ALBERT-base from R/configs/config.yml (hidden 768, 12 shared layers, 12 heads, FFN 2048,
vocab 188), bf16, 192 x 512 synthetic phoneme tokens per GPU.  The batches come from the reference's own input pipeline on
synthetic dataset rows: ``plbert_b200.data.MaskedPhonemeDataset`` + ``Collater`` — mirrors that are
bit-identical to R/dataloader.py's classes (tests/test_data_pipeline.py) — i.e. whole-word 15 % selection, 80/10/10
mask / replace / keep, random 512-crop.
``value`` counts valid phoneme tokens per second over all ranks with the batch already resident in HBM; ``e2e`` is the same
step through the public API from HOST data: dataset rows -> masking + collate in background worker processes
(``CollateWorkerLoader``) -> ``BatchFeeder.step`` (flatten, pinned buffer, H2D, fused step) -> ``FusedAdamW.step`` ->
``loss.item()`` every step.  Both loops time exactly K steps after W warm-up steps in the same process, the end-to-end loop first
(the board drifts ~0.1 ms per 10 steps on its power cap, profiles/e2e_gap.py; ``e2e.order`` records it).
At N = 1 the line also carries: ``sub_records`` (phoneme-only step, C4 variable-length batches in the packed layout, C5
encoder-only forward), ``library_baseline`` (the reference model through stock PyTorch libraries on the SAME GPU: HF
AlbertModel under bf16 autocast + SDPA + torch AdamW, run in a subprocess) with a check of this path's first-step loss
against that run's on the same batch and weights, and ``cpu_baseline`` (the CPU port on the host cores).
``--impl reference`` times the CPU port of the reference step (oracle/reference_port.py) on the host cores.
"""
from __future__ import annotations

import argparse
import itertools
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
DATASET_PARAMS = dict(word_separator=87, max_seq_length=512, word_pred_prob=0.15, phoneme_mask_prob=0.8,
                      replace_prob=0.1)     # R/configs/config.yml:25-30
NUM_PHONEMES = 188   # len(symbols), R/char_indexer.py:15
NUM_TOKENS = 64000   # aubmindlab/bert-base-arabertv2 vocab (R/convert_to_hf.py:24,32) — assumption, no network here
BATCH, SEQ = 192, 512
CPU_SAMPLE_BATCH = 8  # BASELINE.json configs[0]: the reference's own CPU-runnable case (8 x 512)
METRIC = "phoneme_tokens_per_sec_training"
BATCHES_PER_EPOCH = 8


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
    """(bytes, source): average DRAM bytes (read + write) per GEMM launch of one step from the NEWEST committed ncu capture
    of the same step (profiles/r*_gemm_dram_traffic*.json; `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`,
    reduced by profiles/gemm_dram_traffic.py).  A static figure — ncu cannot run inside the bench — so its file is named."""
    import glob
    suffix = "_multitask.json" if two_heads else ".json"
    files = sorted(f for f in glob.glob(os.path.join(ROOT, "profiles", "r*_gemm_dram_traffic*.json"))
                   if f.endswith("_gemm_dram_traffic" + suffix))
    for path in reversed(files):
        try:
            with open(path) as f:
                return json.load(f)["avg_dram_bytes_per_launch"], "static: ncu capture " + os.path.relpath(path, ROOT)
        except (OSError, KeyError, ValueError):
            continue
    return None, None


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


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


def make_loader(rank, two_heads, batch=BATCH, num_workers=0, batches_per_epoch=BATCHES_PER_EPOCH):
    """The benchmark's input pipeline: synthetic dataset rows -> the reference's masking + collate (mirrors), per-rank seed."""
    from plbert_b200 import data as D
    records = D.synthetic_records(batch * batches_per_epoch, seed=1234 + rank, num_tokens=NUM_TOKENS)
    ds = D.MaskedPhonemeDataset(records, use_token_ids=two_heads, **DATASET_PARAMS)
    return D.CollateWorkerLoader(ds, batch, D.Collater() if two_heads else D.PhonemeOnlyCollater(), num_workers=num_workers,
                                 shuffle=True, drop_last=True, seed=1234 + rank)


def make_batch(seed, batch, two_heads):
    """One batch of the pipeline above (first batch of epoch 0 of the loader seeded ``seed``), built in the calling thread."""
    return next(iter(make_loader(seed - 1234, two_heads, batch=batch, batches_per_epoch=1)))


def workload_name(two_heads):
    heads = f"phoneme head V={NUM_PHONEMES} + token head V={NUM_TOKENS} (MultiTaskModel)" if two_heads else \
        f"phoneme head V={NUM_PHONEMES} (PhonemeOnlyModel)"
    return (f"BASELINE.json configs[1]: PL-BERT config.yml ALBERT (hidden 768, 12 shared layers, 12 heads, ffn 2048), {heads}, "
            f"one bf16 train step (fwd + masked CE + bwd + AdamW), {BATCH}x{SEQ} synthetic phoneme tokens per GPU, 15% word masking "
            f"through the reference's dataset + collate pipeline")


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: the reference step (port) on the host cores, bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
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
            "config": {"workload": workload_name(two_heads), "sample": sample, "device": "cpu", "cpu_model": cpu_model()},
            "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample,
                             "cpu_model": cpu_model()},
            "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_library(args):
    """The reference model through stock PyTorch libraries on one B200 (SURVEY.md §8d "PyTorch-library bar"): HF AlbertModel
    under torch.autocast(bf16) with SDPA attention, nn.Linear heads on EVERY position (R/model.py:15-16), the reference's
    per-sample loss loop (R/train.py:107-131), torch.optim.AdamW — same seeded weights and same batch 0 as the b200 arm.
    Informational bar + loss cross-check; not the reference arm (that is the CPU port)."""
    import torch
    from torch import nn
    from transformers import AlbertConfig, AlbertModel

    class Ref(nn.Module):  # R/model.py:5-30
        def __init__(self, enc, two):
            super().__init__()
            self.encoder = enc
            self.phoneme_predictor = nn.Linear(768, NUM_PHONEMES)
            if two:
                self.token_predictor = nn.Linear(768, NUM_TOKENS)

        def forward(self, ids, attention_mask=None):
            h = self.encoder(ids, attention_mask=attention_mask).last_hidden_state
            return [self.phoneme_predictor(h)] + ([self.token_predictor(h)] if hasattr(self, "token_predictor") else [])

    def masked_loss(pred, labels, lengths, midx, crit):  # R/train.py:107-131
        loss, count = 0, 0
        for p, t, n, m in zip(pred, labels, lengths, midx):
            if len(m) > 0:
                loss = loss + crit(p[:n][m], t[:n][m])
                count += 1
        return loss / count

    dev = torch.device("cuda:0")
    two = args.model == "multitask"
    torch.manual_seed(0)
    model = Ref(AlbertModel(AlbertConfig(vocab_size=NUM_PHONEMES, **MODEL_PARAMS)), two).to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=7e-5)
    crit = nn.CrossEntropyLoss()
    hb = next(iter(make_loader(0, two)))
    tok = None
    if two:
        tok, lab, msk, lens, midx = hb
        tok = tok.to(dev)
    else:
        lab, msk, lens, midx = hb
    lab, msk = lab.to(dev), msk.to(dev)

    def step():
        mask = (torch.arange(msk.shape[1], device=dev)[None, :] < torch.tensor(lens, device=dev)[:, None]).int()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            preds = model(msk, attention_mask=mask)
        loss = masked_loss(preds[0].float(), lab, lens, midx, crit)
        if two:
            loss = loss + masked_loss(preds[1].float(), tok, lens, midx, crit)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    first = float(step())          # initial weights, batch 0: the number the b200 arm's first step is checked against
    step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"impl": "library", "what": "reference model through PyTorch libraries on the same B200 (HF AlbertModel, bf16 "
                      "autocast, SDPA, heads on all positions, python loss loop, torch AdamW)", "model": args.model,
                      "ms_per_step": ms, "tokens_per_s": sum(lens) / (ms * 1e-3), "first_step_loss": first, "steps": args.steps,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}), flush=True)


def library_baseline(model_name, steps=3, timeout=600):
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "library", "--model", model_name, "--steps", str(steps)],
                           capture_output=True, text=True, timeout=timeout)
        for line in reversed(r.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return {"unavailable": (r.stderr.strip().splitlines() or ["no output"])[-1][:300]}
    except Exception as e:  # timeout, OOM of the 124 GB all-position logits path, ...
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}


# ------------------------------------------------------------------------------------------------
def _timed(fn, steps, warmup, sync):
    import torch
    for i in range(warmup):
        fn(i)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    sync()
    return e0.elapsed_time(e1) / steps


def sub_records(dev, steps):
    """N = 1 only: the other BASELINE.json configurations that fit one GPU, each through the public API."""
    import torch
    import plbert_b200 as P
    from plbert_b200.batch import synthetic_collate_batch
    from transformers import AlbertConfig, AlbertModel
    out = {}
    sync = lambda: torch.cuda.synchronize(dev)
    torch.manual_seed(0)
    model = P.PhonemeOnlyModel(AlbertModel(AlbertConfig(vocab_size=NUM_PHONEMES, **MODEL_PARAMS)), NUM_PHONEMES, 768).to(dev)
    opt = P.FusedAdamW(model.parameters(), lr=7e-5)
    # ---- phoneme-only step (what R/train.py actually trains), batches resident in HBM ----
    loader = make_loader(0, False)
    resident = [model.stage_batch(b, stager=model.new_stager()) for b in itertools.islice(iter(loader), 2)]

    def po_step(i):
        model.loss_and_backward(resident[i % 2])
        opt.step()

    ms = _timed(po_step, steps, 3, sync)
    fl = step_flops([SEQ] * BATCH, resident[0].M, False)
    pk = measured_peaks()
    out["phoneme_only"] = {"config": "configs[1] with PhonemeOnlyModel (one head, V=188), batches resident in HBM", "ms_per_step": ms,
                           "tokens_per_s": resident[0].num_valid_tokens / (ms * 1e-3), "masked_rows_per_step": resident[0].M,
                           "achieved_tflops": fl / (ms * 1e-3) / 1e12, "vs_burst": fl / (ms * 1e-3) / 1e12 / pk["burst"]}
    # ---- C4: lengths U[32, 512], packed (cu_seqlens) layout, host batch -> step ----
    batches = [synthetic_collate_batch(BATCH, SEQ, seed=77 + i, variable_length=True, min_length=32) for i in range(2)]
    tokens = [int(sum(b[2])) for b in batches]
    for packed in (True, False):
        model.pack_sequences = packed

        def c4_step(i):
            model.loss_and_backward(batches[i % 2])
            opt.step()

        ms = _timed(c4_step, steps, 3, sync)
        tok = sum(tokens[i % 2] for i in range(steps)) / steps
        out["c4_varlen_" + ("packed" if packed else "padded")] = {
            "config": "configs[3]: 192 sequences, lengths U[32,512], phoneme-only step from host batches",
            "ms_per_step": ms, "valid_tokens_per_step": tok, "rows_per_step": model.last_batch.T,
            "valid_tokens_per_s": tok / (ms * 1e-3)}
    model.pack_sequences = True
    # ---- C5: encoder-only forward 64 x 256 (the StyleTTS call, R/README.md:91) through the fused AlbertModel.forward ----
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(1, 185, (64, 256), generator=g).to(dev)
    mask = torch.ones_like(ids, dtype=torch.int32)
    enc = model.encoder
    for use_graph in (False, True):
        enc.plb_use_graph = use_graph

        def c5(i):
            with torch.no_grad():
                return enc(ids, attention_mask=mask).last_hidden_state

        ms = _timed(c5, 50, 5, sync)
        out["c5_encoder" + ("_cuda_graph" if use_graph else "")] = {
            "config": "configs[4]: encoder(ids, attention_mask=m).last_hidden_state, 64x256, bf16 kernels, no grad",
            "ms": ms, "tokens_per_s": 64 * 256 / (ms * 1e-3)}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        hf = lambda i: enc.hf_forward(ids, attention_mask=mask).last_hidden_state   # stock HF forward on the same weights
        ms = _timed(hf, 20, 3, sync)
        a, b = enc(ids, attention_mask=mask).last_hidden_state.float().flatten(), hf(0).float().flatten()
    out["c5_encoder_hf_autocast_bf16"] = {"ms": ms, "tokens_per_s": 64 * 256 / (ms * 1e-3),
                                          "cosine_vs_fused": float(torch.nn.functional.cosine_similarity(a, b, dim=0))}
    del model, opt, resident
    return out


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
    from plbert_b200.batch import BatchFeeder
    from plbert_b200.ddp import GradientSynchronizer
    from transformers import AlbertConfig, AlbertModel

    two_heads = args.model == "multitask"
    torch.manual_seed(0)
    enc = AlbertModel(AlbertConfig(vocab_size=NUM_PHONEMES, **MODEL_PARAMS))
    model = (P.MultiTaskModel(enc, NUM_PHONEMES, NUM_TOKENS, 768) if two_heads else P.PhonemeOnlyModel(enc, NUM_PHONEMES, 768)).to(dev)
    opt = P.FusedAdamW(model.parameters(), lr=7e-5)
    heads = ["phoneme_predictor", "token_predictor"] if two_heads else ["phoneme_predictor"]

    # the input pipeline: worker processes build the batches (forked BEFORE any CUDA work of this process)
    cores = os.cpu_count() or 8
    workers = max(2, min(12, cores // max(world, 1) - 2))
    loader = make_loader(rank, two_heads, num_workers=workers)

    stream = loader.endless()   # all epochs back to back: the workers never drain
    n_batches = 2
    host_batches = [next(stream) for _ in range(n_batches)]     # epoch 0, batches 0 and 1 (batch 0 is the loss-check batch)
    eng = model._ensure_engine(dev)
    sync = GradientSynchronizer(eng) if world > 1 else None
    if sync is not None:
        opt.grad_scale = sync.grad_scale
    lr, gs = 7e-5, opt.grad_scale

    # resident batches for the device-only number (their own stager slots so they are not recycled)
    resident = [model.stage_batch(hb, stager=model.new_stager()) for hb in host_batches]
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

    eng.loss_and_backward(resident[0], heads)      # first step: initial weights, batch 0 (the library arm repeats exactly this)
    first_loss = float(eng.loss_buf[0].item())
    eng.adamw_step(lr, grad_scale=gs)
    for i in range(1, args.warmup):
        device_step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # ---- end to end through the public API: dataset rows -> worker processes (masking + collate) -> BatchFeeder (flatten,
    # pinned buffer, H2D, step) -> optimizer -> loss.item(); every step consumes a NEW batch from the workers.  Timed BEFORE
    # the device-resident loop: the board sits on its power cap and loses ~0.1 ms per step for every 10 steps it has been
    # running (profiles/e2e_gap.py: resident 53.9 -> 54.7 -> 55.0 ms, end to end 54.5 -> 54.8 ms in alternation), so whichever
    # loop runs second looks ~2 % slower; the end-to-end number is the headline, so it gets the same (cooler) board the
    # device-resident number had in round 1 ----
    feed = BatchFeeder(model, stream)
    e2e_tokens = [0]

    def api_step(i):
        n = feed._staged.num_valid_tokens
        loss = feed.step()
        opt.step()
        e2e_tokens[0] += n
        return loss.item()

    for i in range(2):
        api_step(i)
    barrier()
    e2e_tokens[0] = 0
    e0.record()
    for i in range(args.steps):
        last_loss = api_step(i)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    tk = torch.tensor([float(e2e_tokens[0])], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tk, op=dist.ReduceOp.SUM)
    e2e_ms = t.item() / args.steps
    e2e_value = tk.item() / (t.item() * 1e-3)
    h2d = model.last_batch.h2d_bytes

    # ---- device-resident loop: batches already in HBM ----
    barrier()
    l0 = eng.launches
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
        traffic, traffic_src = gemm_dram_traffic(two_heads)
        roof = {"bound": "tensor", "kernel": "gemm_kernel (tcgen05/TMA; fwd+dgrad+wgrad, all epilogues)", "achieved": achieved,
                "peak": pk["sustained"], "unit": "TFLOP/s", "frac": achieved / pk["sustained"],
                "traffic": traffic, "traffic_source": traffic_src,
                "launches_per_step": len(records), "avg_launch_us": tt / len(records) * 1e6,
                "flops_per_launch": fl / len(records), "gemm_share_of_step": tt / (ms_total * 1e-3 / args.steps),
                "peak_source": pk["source"] + ", sustained figure (kernel timed inside a long step); burst " + str(pk["burst"])}
    if world > 1:
        barrier()

    line = None
    if rank == 0:
        pk = measured_peaks()
        fl_step = sum(step_flops([SEQ] * BATCH, rows_per_step[i % n_batches], two_heads) for i in range(args.steps)) / args.steps
        step_s = ms_total * 1e-3 / args.steps
        line = {"metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": workload_name(two_heads), "parallelism": f"dp{world}",
                           "l2": "inputs larger than L2: each step streams ~24 GB of bf16 activations (126 MB L2), two batches alternate",
                           "tokens_per_step_per_gpu": tokens_per_step[0], "masked_rows_per_step": rows_per_step[0],
                           "first_step_loss": first_loss, "loss": last_loss, "input_workers": workers},
                "per_gpu": value / world,
                "tensor_pipe_frac_of_bf16_peak": {"algorithmic_tflop_per_step": fl_step / 1e12,
                                                  "achieved_tflops": fl_step / step_s / 1e12,
                                                  "vs_burst": fl_step / step_s / 1e12 / pk["burst"],
                                                  "vs_sustained": fl_step / step_s / 1e12 / pk["sustained"]},
                "roofline": roof, "cpu_baseline": None,
                "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": e2e_ms, "order": "timed before the device-resident loop (same process, same warm-up)",
                        "pipeline": "dataset rows -> masking + collate in worker processes -> pinned H2D -> step -> loss.item()"},
                "gpu_launches": launches, "clocks": clocks}
    if world > 1:
        dist.destroy_process_group()
    if rank != 0 or world > 1:
        if line is not None:
            print(json.dumps(line), flush=True)
        return

    # ---------------- N = 1 extras: sub-records, library bar (+ loss check), CPU baseline ----------------
    import gc
    del feed, resident, opt, model, eng, enc, sync, stream, loader   # frees ~30 GB and stops the input workers
    gc.collect()                                                      # encoder <-> owner model reference cycle (fuse_albert)
    torch.cuda.empty_cache()
    if not args.no_sub_records:
        line["sub_records"] = sub_records(dev, max(4, min(args.steps, 8)))
        gc.collect()
        torch.cuda.empty_cache()
    ok = True
    if not args.no_library:
        lib = library_baseline(args.model)
        line["library_baseline"] = lib
        if "first_step_loss" in lib:
            rel = abs(first_loss - lib["first_step_loss"]) / abs(lib["first_step_loss"])
            ok = rel <= 2e-2
            line["loss_check"] = {"this_path_first_step": first_loss, "library_first_step": lib["first_step_loss"], "rel_err": rel,
                                  "bar": 2e-2, "pass": ok, "what": "same seeded weights, same batch 0 (192x512), bf16 both"}
            line["library_baseline"]["speedup_of_this_path"] = lib["ms_per_step"] / line["ms_per_step"]
    if not args.no_cpu_baseline:
        from oracle import reference_port as RP
        cb = make_batch(1234, CPU_SAMPLE_BATCH, two_heads)
        dt, ctok, threads = RP.time_steps(cb, NUM_TOKENS if two_heads else None, 2, 1)
        line["cpu_baseline"] = {"value": ctok / dt, "unit": "tokens/s", "cores": threads, "kind": "port", "cpu_model": cpu_model(),
                                "sample": f"{CPU_SAMPLE_BATCH}x{SEQ} tokens per step, fp32, 1 warm-up + 2 timed steps of the reference step "
                                          f"(oracle/reference_port.py: HF AlbertModel + heads + loss loop + torch AdamW)"}
    print(json.dumps(line), flush=True)
    if not ok:
        raise SystemExit("bench: first-step loss differs from the PyTorch-library run on the same batch by more than 2e-2")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "library"])
    ap.add_argument("--model", default="multitask", choices=["multitask", "phoneme_only"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library", action="store_true")
    ap.add_argument("--no-sub-records", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "library":
        run_library(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_b200(args)


if __name__ == "__main__":
    main()
