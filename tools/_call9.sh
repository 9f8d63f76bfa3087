set -x
timeout 900 python tools/gpu_conv_sweep.py > gpurun_out/sweep_v11_full.log 2>&1; cp gpurun_out/conv_sweep.md gpurun_out/conv_sweep_v11_full.md
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_v11.csv python bench.py --profile-step > gpurun_out/ncu_list11.log 2>&1
tail -2 gpurun_out/ncu_list11.log
