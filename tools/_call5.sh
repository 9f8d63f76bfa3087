set -x
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu_c5.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_c5.log
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_v9_two.json 2> gpurun_out/bench_v9_two.err
tail -3 gpurun_out/bench_v9_two.err; cut -c1-300 gpurun_out/bench_v9_two.json
W2C_TWO_STREAMS=0 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_v9_one.json 2> gpurun_out/bench_v9_one.err
cut -c1-300 gpurun_out/bench_v9_one.json
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_v9_two_b.json 2> gpurun_out/bench_v9_two_b.err
cut -c1-300 gpurun_out/bench_v9_two_b.json
