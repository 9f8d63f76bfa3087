set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "squeeze or selection or comnet or additive or normal_agents" > gpurun_out/pytest_gpu_c16.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_gpu_c16.log | cut -c1-300
