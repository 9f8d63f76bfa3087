set -x
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_glue.py -m gpu -q -x > gpurun_out/pytest_gpu_c19.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_c19.log | cut -c1-300
SWEEP_DIRECT=1 timeout 600 python tools/gpu_conv_sweep.py > gpurun_out/sweep_direct.log 2>&1; cp gpurun_out/conv_sweep.md gpurun_out/conv_sweep_direct.md
for cfgs in "2 1" "0 0" "2 0" "0 1" "1 1"; do set -- $cfgs
W2C_CONV_DIRECT=$1 W2C_STEM_DIRECT=$2 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-fused-e2e --no-parity-value > gpurun_out/bench_direct_$1_$2.json 2> gpurun_out/bench_direct_$1_$2.err
cut -c1-200 gpurun_out/bench_direct_$1_$2.json; tail -2 gpurun_out/bench_direct_$1_$2.err
done
