#!/bin/bash
# Soak run of the training driver at the benchmarked size on ONE B200: config.yml ALBERT (hidden 768, 12 shared layers), two heads
# (V 188 + 64 000), 192 x 512 per step, synthetic dataset rows through the reference's masking + collate pipeline in worker
# processes, 300 optimizer steps, validation + checkpoint at step 300.  Prints the logged losses (finite, phoneme loss falling).
set -e
mkdir -p /tmp/soak && cd /root/repo
python - <<'PY'
import yaml
cfg = {"training_params": dict(output_dir="/tmp/soak/runs", mixed_precision="fp16", batch_size=192, save_interval=300, log_interval=25,
                               num_steps=300, learning_rate=7e-5, training_dataset="synthetic_records", split="x",
                               synthetic_rows=8192),
       "dataset_params": dict(word_separator=87, max_seq_length=512, word_pred_prob=0.15, phoneme_mask_prob=0.8, replace_prob=0.1),
       "model_params": dict(pretrained_model="", hidden_size=768, num_attention_heads=12, intermediate_size=2048,
                            max_position_embeddings=512, num_hidden_layers=12, dropout=0.1),
       "b200_params": dict(use_token_ids=True, num_tokens=64000, num_workers=10, log_lag=1)}
yaml.dump(cfg, open("/tmp/soak/config.yml", "w"))
PY
python -m plbert_b200.train --config_path /tmp/soak/config.yml --run_name soak --no_wandb 2>&1 | grep -v Warning | tail -6
ls -la /tmp/soak/runs/soak | tail -4
cat /tmp/soak/runs/soak/metrics.jsonl
