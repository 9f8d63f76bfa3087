set -x
export SWEEP_PROD=1
for v in electtop uniform electtop uniform; do
  W2C_LIB=$PWD/multiagentperception_b200/lib/libw2c_$v.so timeout 600 python tools/gpu_conv_sweep.py > gpurun_out/sweep_$v.log 2>&1
  cp gpurun_out/conv_sweep.md gpurun_out/conv_sweep_ab_$v.md
  grep -c ERROR gpurun_out/conv_sweep_ab_$v.md
  mv gpurun_out/conv_sweep_ab_$v.md gpurun_out/conv_sweep_ab_${v}_$(date +%s).md
done
for v in electtop uniform; do
W2C_LIB=$PWD/multiagentperception_b200/lib/libw2c_$v.so python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-fused-e2e --no-parity-value > gpurun_out/bench_ab_$v.json 2> gpurun_out/bench_ab_$v.err
cut -c1-200 gpurun_out/bench_ab_$v.json
done
