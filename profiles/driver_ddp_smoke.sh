set -e
mkdir -p /tmp/drv && cd /root/repo
python - <<'PY'
import sys, yaml
sys.path.insert(0, "tests")
from test_driver_cpu import write_config
write_config("/tmp/drv/config.yml", "/tmp/drv/runs", synthetic_batches_per_epoch=4)
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 -m plbert_b200.train --config_path /tmp/drv/config.yml --run_name ddp2 --no_wandb
ls /tmp/drv/runs/ddp2; cat /tmp/drv/runs/ddp2/metrics.jsonl | tail -4
