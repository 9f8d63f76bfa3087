set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_c15.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu_c15.log | cut -c1-300
grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_c15.log | head -20
