set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_c27.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_c27.log | cut -c1-300
grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_c27.log | head
python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo "bench rc=$?"; wc -l gpurun_out/bench_final2.json
python -c "
import json;d=json.load(open('gpurun_out/bench_final2.json'));print('final',round(d['value']),round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'fused',round(d['e2e_fused']['value']),'x3',round(d['parity_precision']['value']),'fp16',round(d['fp16_precision']['value']),'conv',round(d['roofline']['achieved']),round(d['roofline']['frac'],3),d['clocks']['sm_mhz'],d['gpu_launches'])"
