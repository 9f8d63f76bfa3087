set -x
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu_c12.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_c12.log
SWEEP_PF=1 timeout 600 python tools/gpu_conv_sweep.py > gpurun_out/sweep_pf.log 2>&1; cp gpurun_out/conv_sweep.md gpurun_out/conv_sweep_pf.md
for pf in 2 0 1 3; do
W2C_CONV_PREFETCH=$pf python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-fused-e2e --no-parity-value > gpurun_out/bench_pf$pf.json 2> gpurun_out/bench_pf$pf.err
cut -c1-200 gpurun_out/bench_pf$pf.json
done
