#!/bin/bash
# round-2 GPU session G: new ASG (chains + recompute) — parity, memcheck, timing
mkdir -p gpurun_out && rm -f gpurun_out/arch_parity.jsonl
timeout 900 python -m pytest tests/test_gpu_criterion.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/g_pytest_crit.log 2>&1; echo "pytest exit $?" >> gpurun_out/g_pytest_crit.log
tail -25 gpurun_out/g_pytest_crit.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_criterion.py -m gpu -q --tb=line -p no:cacheprovider -k "test_asg_parity and not baseline" > gpurun_out/g_memcheck.log 2>&1; tail -12 gpurun_out/g_memcheck.log
timeout 300 python bench.py --workload asg --steps 20 --warmup 3 --no-cpu > gpurun_out/g_bench_asg.json 2> gpurun_out/g_bench_asg.err; tail -c 1500 gpurun_out/g_bench_asg.json; tail -3 gpurun_out/g_bench_asg.err
timeout 900 python -m pytest tests/test_gpu_archs.py tests/test_gpu_export.py tests/test_gpu_trainer.py tests/test_gpu_convglu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/g_pytest_rest.log 2>&1; tail -15 gpurun_out/g_pytest_rest.log
grep -o '"arch": "[a-z_0-9]*", "precision": "f32".\{0,60\}\|"scalar_ln_backward_error": [0-9.e-]*' gpurun_out/arch_parity.jsonl | head -20
