set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_c3.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu_c3.log
python bench.py --steps 20 --warmup 5 --layer-table gpurun_out/r1_layers_v7.md > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err
tail -3 gpurun_out/bench_v7.err; cut -c1-300 gpurun_out/bench_v7.json
W2C_STEM_STAGING=2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layer-table gpurun_out/r1_layers_v7_ns2.md > gpurun_out/bench_v7_ns2.json 2> gpurun_out/bench_v7_ns2.err
head -4 gpurun_out/r1_layers_v7.md gpurun_out/r1_layers_v7_ns2.md
SWEEP_AB=1 timeout 600 python tools/gpu_conv_sweep.py > gpurun_out/sweep_ab2.log 2>&1; cp gpurun_out/conv_sweep.md gpurun_out/conv_sweep_ab2.md
