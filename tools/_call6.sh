set -x
export SWEEP_PROD=1 SWEEP_ONLY="1,512,64,64;2,256,64,64;0,512,64,11;0,256,128,64"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_persv1 --launch-skip 6 --launch-count 1 -o gpurun_out/narrow_a python tools/gpu_conv_sweep.py > gpurun_out/ncu_narrow_a.log 2>&1
export SWEEP_ONLY="2,256,64,64"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_persv1 --launch-skip 6 --launch-count 1 -o gpurun_out/narrow_b python tools/gpu_conv_sweep.py > gpurun_out/ncu_narrow_b.log 2>&1
export SWEEP_ONLY="0,512,64,11"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_persv1 --launch-skip 6 --launch-count 1 -o gpurun_out/narrow_c python tools/gpu_conv_sweep.py > gpurun_out/ncu_narrow_c.log 2>&1
export SWEEP_ONLY="0,256,128,64"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_persv1 --launch-skip 6 --launch-count 1 -o gpurun_out/narrow_d python tools/gpu_conv_sweep.py > gpurun_out/ncu_narrow_d.log 2>&1
ls -la gpurun_out/*.ncu-rep
