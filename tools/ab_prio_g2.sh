run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/s4_g2_$name.log 2>&1
  echo "$name rgb: $(tail -1 gpurun_out/s4_g2_$name.log | grep -o '"ms_per_step": [0-9.]*')"
}
run a SANERF_CRIT_PRIO=-1 SANERF_UPD_PRIO=0
run b SANERF_CRIT_PRIO=0 SANERF_UPD_PRIO=0
run c SANERF_CRIT_PRIO=-1 SANERF_UPD_PRIO=-1
run d SANERF_CRIT_PRIO=0 SANERF_UPD_PRIO=-1
