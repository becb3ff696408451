#!/bin/bash
# round-2 GPU session W: full GPU suite + the driver's default bench invocation after the ASG rewrite
mkdir -p gpurun_out && rm -f gpurun_out/arch_parity.jsonl
( time timeout 1500 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider ) > gpurun_out/w_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/w_pytest.log
tail -12 gpurun_out/w_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/w_bench_default.json 2> gpurun_out/w_bench_default.err; tail -5 gpurun_out/w_bench_default.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/w_bench_reference.json 2> gpurun_out/w_bench_reference.err; tail -4 gpurun_out/w_bench_reference.err; tail -c 600 gpurun_out/w_bench_reference.json
python __graft_entry__.py smoke 2>&1 | tail -2
