#!/usr/bin/env bash
# ncu passes of one bench step (run on a B200 under gpurun; everything lands in gpurun_out/):
#   1. launch list with DRAM traffic  -> tools/summarize_ncu_launches.py -> profiles/r2_ncu_launches_vN.md + r2_traffic.json
#   2. --set full of the dominant kernel (the 512->512 @64x64 persistent conv: the 7th conv_persv1 launch of the step now that the encoder head is one fused kernel)
# Numbers printed by a run under ncu are never bench values.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum \
    --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --profile-step > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_persv1 --launch-skip 6 --launch-count 1 \
    -o gpurun_out/top_kernel python bench.py --profile-step > gpurun_out/ncu_top.log 2>&1
ncu -i gpurun_out/top_kernel.ncu-rep --page raw --csv > gpurun_out/top_kernel_raw.csv 2>/dev/null
ls -la gpurun_out/
