#!/bin/bash
# round-2 GPU session D: conv path comparison, full tests, benches per precision
mkdir -p gpurun_out && rm -f gpurun_out/arch_parity.jsonl
timeout 300 python scripts/bench_conv_paths.py > gpurun_out/d_conv_paths.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/d_pytest.log
tail -5 gpurun_out/d_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/d_bench_tds.json 2> gpurun_out/d_bench_tds.err
timeout 300 python bench.py --precision tf32 --steps 10 --warmup 3 --no-extras --no-cpu > gpurun_out/d_bench_tds_tf32.json 2> gpurun_out/d_bench_tds_tf32.err
timeout 300 python bench.py --precision bf16 --steps 10 --warmup 3 --no-extras --no-cpu > gpurun_out/d_bench_tds_bf16.json 2> gpurun_out/d_bench_tds_bf16.err
timeout 300 python bench.py --workload streaming_tds_ctc --steps 5 --warmup 5 --no-cpu > gpurun_out/d_bench_streaming.json 2> gpurun_out/d_bench_streaming.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/d_bench_ref.json 2> gpurun_out/d_bench_ref.err
timeout 300 ncu --set full --clock-control none -k regex:conv_umma_fwd -c 3 -o /tmp/d_prof_conv python scripts/prof_conv.py > gpurun_out/d_ncu_conv.log 2>&1
ncu -i /tmp/d_prof_conv.ncu-rep --page raw --csv > gpurun_out/d_prof_conv_raw.csv 2>/dev/null
du -sh gpurun_out
