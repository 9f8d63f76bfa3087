start=$(date +%s)
python bench.py > gpurun_out/s3_bench_final.json 2> gpurun_out/s3_bench_final.err
echo "bench rc=$? wall=$(( $(date +%s) - start ))s" >> gpurun_out/s3_bench_final.err
