python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/gpu_shard_check.py > gpurun_out/shard_check_2gpu_final.log 2>&1; echo "shard rc=$?"
grep -c '"sharded_equals_unsharded": true' gpurun_out/shard_check_2gpu_final.log; grep -c '"sharded_equals_unsharded": false' gpurun_out/shard_check_2gpu_final.log
