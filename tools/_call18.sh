set -x
SWEEP_ONLY="2,256,64,64" timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_persv1 --launch-skip 27 --launch-count 1 -o gpurun_out/dense_deconv python tools/gpu_conv_sweep.py > gpurun_out/ncu_dense.log 2>&1
ls -la gpurun_out/dense_deconv.ncu-rep
