#!/bin/bash
# round-2 GPU session A: full GPU test suite, GEMM kinds microbench, benches of every workload, launch list
mkdir -p gpurun_out && rm -f gpurun_out/arch_parity.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/a_smi.txt 2>&1
nproc > gpurun_out/a_nproc.txt; lscpu | grep "Model name" >> gpurun_out/a_nproc.txt
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 300 python scripts/bench_gemm_kinds.py > gpurun_out/a_gemm_kinds.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench_tds.json 2> gpurun_out/a_bench_tds.err
timeout 300 python bench.py --workload conv_glu_asg --steps 5 --no-cpu > gpurun_out/a_bench_convglu.json 2> gpurun_out/a_bench_convglu.err
timeout 300 python bench.py --workload streaming_tds_ctc --steps 5 --no-cpu > gpurun_out/a_bench_streaming.json 2> gpurun_out/a_bench_streaming.err
timeout 300 python bench.py --workload asg --steps 20 > gpurun_out/a_bench_asg.json 2> gpurun_out/a_bench_asg.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/a_launches_tds_f32.csv \
   python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > gpurun_out/a_ncu_bench.log 2>&1
ls -la gpurun_out | head -40
