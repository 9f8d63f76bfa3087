set -x
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu_c10.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_c10.log
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-fused-e2e --no-parity-value --layer-table gpurun_out/r1_layers_v12.md > gpurun_out/bench_v12.json 2> gpurun_out/bench_v12.err
tail -3 gpurun_out/bench_v12.err; cut -c1-200 gpurun_out/bench_v12.json
export SWEEP_PROD=1
for spec in "a2:1,512,64,64" "b2:2,256,64,64" "c2:0,512,64,11" "e2:0,256,64,128"; do
  name=${spec%%:*}; export SWEEP_ONLY=${spec#*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_persv1 --launch-skip 6 --launch-count 1 -o gpurun_out/narrow_$name python tools/gpu_conv_sweep.py > gpurun_out/ncu_narrow_$name.log 2>&1
done
ls -la gpurun_out/*2.ncu-rep
