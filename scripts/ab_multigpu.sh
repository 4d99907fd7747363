for cfg in "B2U_STATIC_TILES=1" "B2U_STATIC_TILES=0" "B2U_STATIC_TILES=0 NCCL_MAX_CTAS=8" "B2U_STATIC_TILES=0 B2U_BUCKET_MB=32"; do
  env $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $1 --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', 'N=', d['n_gpus'], round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms  dp_diff', d.get('dp_param_max_diff'))"
done
B2U_STATIC_TILES=0 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1', round(d['value'],1), round(d['ms_per_step'],3))"
