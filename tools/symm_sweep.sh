# usage: bash tools/symm_sweep.sh <gpus> "<threads> <blocks> <p2p|multimem>" ...
N=$1; shift
for cfg in "$@"; do set -- $cfg
SANERF_SYMM_THREADS=$1 SANERF_SYMM_BLOCKS=$2 SANERF_SYMM_REDUCE=$3 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --workload rgb --no-e2e --no-grad-equiv 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('gpus $N threads $1 blocks $2 reduce $3', round(d['ms_per_step'],4), round(d['ms_per_step_median'],4), [ (k,v) for k,v in d['launch_table_ms'].items() if 'symm' in k])"
done
if [ -n "$NCCL_BASELINE" ]; then
SANERF_SYMM=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --workload rgb --no-e2e --no-grad-equiv 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('gpus $N NCCL exchange', round(d['ms_per_step'],4), round(d['ms_per_step_median'],4))"
fi
