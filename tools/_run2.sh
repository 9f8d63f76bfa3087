python -m pytest tests/test_backward_kernels_gpu.py -x -q -m gpu > gpurun_out/s3_bwdk.log 2>&1
python tools/gpu_train_bench.py --backbones resnet --no-library --precisions bf16 > gpurun_out/s3_train_resnet2.json 2> gpurun_out/s3_train_resnet2.err
python tools/gpu_train_bench.py --no-library --precisions bf16 > gpurun_out/s3_train_nseg2.json 2> gpurun_out/s3_train_nseg2.err
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/s3_train_launches_nseg.csv python tools/profile_train_step.py bf16 n_segnet > gpurun_out/s3_prof2.log 2>&1
