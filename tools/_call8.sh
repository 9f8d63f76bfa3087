set -x
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_glue.py tests/test_parity_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu_c8.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_c8.log
SWEEP_ONLY="0,512,64,11" timeout 600 python tools/gpu_conv_sweep.py > gpurun_out/sweep_logits.log 2>&1; cat gpurun_out/conv_sweep.md
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_v11.json 2> gpurun_out/bench_v11.err
tail -3 gpurun_out/bench_v11.err; cut -c1-200 gpurun_out/bench_v11.json
W2C_CONV_NCHW_TMA=0 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-fused-e2e --no-parity-value > gpurun_out/bench_v11_direct.json 2> gpurun_out/bench_v11_direct.err
cut -c1-200 gpurun_out/bench_v11_direct.json
