#!/bin/bash
# quick kernel timing on the GPU box: prints kernel_ms / TFLOP/s of the two hot kernels
python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --slab-steps ${1:-73} 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line)
        print('value %.4e  step %.2f ms  step_tf %.2f  eddy %.2f TF (%.2f ms)  project %.2f TF (%.2f ms)' % (d['value'], d['ms_per_step'], d['step_fp64_tflops_per_gpu'], d['roofline']['achieved'], d['kernel_ms']['eddy_flux_project'], d['roofline_project']['achieved'], d['kernel_ms']['project']))
    else: print(line)
"
