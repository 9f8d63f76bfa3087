set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_c4.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu_c4.log
python bench.py --steps 20 --warmup 5 --layer-table gpurun_out/r1_layers_v8.md > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err
tail -5 gpurun_out/bench_v8.err; cut -c1-300 gpurun_out/bench_v8.json
SWEEP_AB=1 timeout 600 python tools/gpu_conv_sweep.py > gpurun_out/sweep_ab3.log 2>&1; cp gpurun_out/conv_sweep.md gpurun_out/conv_sweep_ab3.md
