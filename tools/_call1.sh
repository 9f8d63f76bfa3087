set -x
python bench.py --steps 20 --warmup 5 --layer-table gpurun_out/r1_layers_v5.md > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err
tail -3 gpurun_out/bench_v5.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_v5.csv python bench.py --profile-step > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off -o gpurun_out/step_full python bench.py --profile-step > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/
ncu -i gpurun_out/step_full.ncu-rep --page raw --csv > gpurun_out/step_full_raw.csv 2>/dev/null
sz=$(stat -c %s gpurun_out/step_full.ncu-rep); if [ "$sz" -gt 45000000 ]; then rm gpurun_out/step_full.ncu-rep; fi
