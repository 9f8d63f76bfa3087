set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_c14.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu_c14.log | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/gpu_shard_check.py > gpurun_out/shard_check_2gpu.log 2>&1; echo "shard rc=$?"
grep sharded_equals gpurun_out/shard_check_2gpu.log | cut -c1-220
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-fused-e2e --no-parity-value > gpurun_out/bench_v13.json 2> gpurun_out/bench_v13.err
cut -c1-200 gpurun_out/bench_v13.json
