set -x
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_c7.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_c7.log
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --layer-table gpurun_out/r1_layers_v10.md > gpurun_out/bench_v10.json 2> gpurun_out/bench_v10.err
tail -3 gpurun_out/bench_v10.err; cut -c1-300 gpurun_out/bench_v10.json
SWEEP_QUICK=1 timeout 900 python tools/gpu_conv_sweep.py > gpurun_out/sweep_v10.log 2>&1; cp gpurun_out/conv_sweep.md gpurun_out/conv_sweep_v10.md
