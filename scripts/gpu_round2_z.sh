#!/bin/bash
# round-2 GPU session Z: block-staged CTC chains, FAC grad at 4 CTAs/SM: parity, memcheck, timing
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_criterion.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/z_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/z_pytest.log
tail -6 gpurun_out/z_pytest.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_criterion.py -m gpu -q --tb=line -p no:cacheprovider -k "ctc or (test_asg_parity and not baseline)" > gpurun_out/z_memcheck.log 2>&1; tail -3 gpurun_out/z_memcheck.log
timeout 300 python bench.py --precision tf32 --steps 10 --warmup 3 --no-extras --no-cpu 2>gpurun_out/z_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('tf32', d['ms_per_step'], d['roofline']['gemm_ms_per_step'])
print({k:(v['ms'],v['launches']) for k,v in list(d['step_breakdown']['kernels'].items()) if 'ctc' in k})"
true
true
true
tail -3 gpurun_out/z_bench.err
