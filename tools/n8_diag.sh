run() { echo "== $1"; env $1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e --no-parity --no-c4 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],3), 'train', round(d['train_ms_per_step'],3), 'rollout', round(d['ms_per_step']-d['train_ms_per_step'],3))"; }
run "X=1"
run "PPO_DISABLE_SHUFFLE_PREFETCH=1"
run "PPO_DISABLE_SHUFFLE_SHARDING=1"
