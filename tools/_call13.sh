set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/gpu_shard_check.py > gpurun_out/shard_check_2gpu.log 2>&1; echo "shard rc=$?"
tail -12 gpurun_out/shard_check_2gpu.log | cut -c1-220
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_2gpu.err; cut -c1-300 gpurun_out/bench_2gpu.json
