#!/bin/bash
# round-2 GPU session F: state of the tree after the container was re-created (tests, all bench workloads)
mkdir -p gpurun_out && rm -f gpurun_out/arch_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/f_pytest.log
tail -5 gpurun_out/f_pytest.log
for p in f32 tf32 bf16; do
  timeout 300 python bench.py --precision $p --steps 10 --warmup 3 --no-cpu > gpurun_out/f_bench_tds_$p.json 2> gpurun_out/f_bench_tds_$p.err
done
timeout 300 python bench.py --workload asg --steps 20 --warmup 3 --no-cpu > gpurun_out/f_bench_asg.json 2> gpurun_out/f_bench_asg.err
timeout 300 python bench.py --workload conv_glu_asg --steps 5 --warmup 8 --no-cpu > gpurun_out/f_bench_convglu.json 2> gpurun_out/f_bench_convglu.err
timeout 300 python bench.py --workload streaming_tds_ctc --steps 5 --warmup 5 --no-cpu > gpurun_out/f_bench_streaming.json 2> gpurun_out/f_bench_streaming.err
tail -c 600 gpurun_out/f_bench_*.json
du -sh gpurun_out
