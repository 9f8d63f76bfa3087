#!/bin/bash
# One 8-GPU gpurun session: the default config at N = 8, the latency config (3) and the batch sweep of config 5.
# Every step has its own timeout and writes straight into gpurun_out/ (merged back even if a later step is killed).
N=${1:-8}
run() {  # name, extra args
  local name=$1; shift
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --steps 20 --warmup 5 "$@" \
      > gpurun_out/r2_n${N}_$name.json 2> gpurun_out/r2_n${N}_$name.err
  echo "$name rc=$? $(cut -c1-160 gpurun_out/r2_n${N}_$name.json)"
}
run cfg2 --no-parity-value
run cfg3 --config 3 --lean
for b in 1 2 4 8 16 32; do run cfg5_b$b --config 5 --batch $b --lean; done
run cfg2_resnet --backbones resnet --lean
