# timing diagnostics at N GPUs: where the data-parallel exchange is exposed (variants 2-3 compute wrong gradients)
N=${1:-8}
run() { name=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/s4_g${N}_$name.log 2>&1
  echo "$name: $(tail -1 gpurun_out/s4_g${N}_$name.log | grep -o '"ms_per_step": [0-9.]*' | head -1)"; }
run full X=1
run nomain SANERF_DBG_SKIP_MAIN_NCCL=1
run notail SANERF_DBG_SKIP_TAIL_NCCL=1
run none SANERF_DBG_SKIP_MAIN_NCCL=1 SANERF_DBG_SKIP_TAIL_NCCL=1
