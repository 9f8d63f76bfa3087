set -x
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "dense or deconv" > gpurun_out/pytest_gpu_c17.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu_c17.log | cut -c1-300
SWEEP_QUICK=0 SWEEP_ONLY="2,256,64,64" timeout 300 python tools/gpu_conv_sweep.py > gpurun_out/sweep_dense.log 2>&1; cat gpurun_out/conv_sweep.md
for d in 1 0; do
W2C_DECONV_DENSE=$d python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-fused-e2e --no-parity-value > gpurun_out/bench_dense$d.json 2> gpurun_out/bench_dense$d.err
cut -c1-200 gpurun_out/bench_dense$d.json; tail -2 gpurun_out/bench_dense$d.err
done
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu_c17b.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/pytest_gpu_c17b.log | cut -c1-200
