SWEEP_ROT=1 timeout 600 python tools/gpu_conv_sweep.py > gpurun_out/sweep_rot.log 2>&1; cat gpurun_out/conv_sweep.md
