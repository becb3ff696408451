#!/bin/bash
# round-2 GPU session Y: lean CTC chains + f32x3 GEMM with eight split warps: parity, memcheck, timing
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_criterion.py tests/test_gpu_gemm.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/y_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/y_pytest.log
tail -15 gpurun_out/y_pytest.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_criterion.py -m gpu -q --tb=line -p no:cacheprovider -k "ctc" > gpurun_out/y_memcheck.log 2>&1; tail -3 gpurun_out/y_memcheck.log
timeout 300 python bench.py --precision f32 --steps 10 --warmup 3 --no-extras --no-cpu 2>gpurun_out/y_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('f32', d['ms_per_step'], d['roofline']['gemm_ms_per_step'], d['roofline']['achieved'])
print({k:(v['ms'],v['launches']) for k,v in list(d['step_breakdown']['kernels'].items())[:14]})"
timeout 300 python bench.py --precision tf32 --steps 10 --warmup 3 --no-extras --no-cpu 2>>gpurun_out/y_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('tf32', d['ms_per_step'], d['roofline']['gemm_ms_per_step'], d['roofline']['achieved'])
print({k:(v['ms'],v['launches']) for k,v in list(d['step_breakdown']['kernels'].items()) if 'ctc' in k})"
tail -3 gpurun_out/y_bench.err
