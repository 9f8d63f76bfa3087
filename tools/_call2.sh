set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c2.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu_c2.log
SWEEP_AB=1 timeout 600 python tools/gpu_conv_sweep.py > gpurun_out/sweep_ab.log 2>&1; cp gpurun_out/conv_sweep.md gpurun_out/conv_sweep_ab.md
python bench.py --steps 20 --warmup 5 --layer-table gpurun_out/r1_layers_v6.md > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err
tail -3 gpurun_out/bench_v6.err; cat gpurun_out/bench_v6.json | cut -c1-400
