set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_c23.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_c23.log | cut -c1-300
python bench.py --backbones resnet --no-cpu-baseline --no-parity-value --layer-table gpurun_out/layers_resnet3.md > gpurun_out/bench_resnet4.json 2>gpurun_out/bench_resnet4.err
python -c "
import json;d=json.load(open('gpurun_out/bench_resnet4.json'));print('resnet',round(d['value']),round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'fused',d['e2e_fused'] and round(d['e2e_fused']['value']))"
grep -E "stem|maxpool|bilinear|mlp|attn|total" gpurun_out/layers_resnet3.md
