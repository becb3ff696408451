#!/bin/bash
# round-2 GPU session C: tcgen05 conv, vector epilogue, f32 diagnosis
mkdir -p gpurun_out && rm -f gpurun_out/arch_parity.jsonl
timeout 300 python scripts/diag_f32.py > gpurun_out/c_diag_f32.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c_pytest.log
tail -5 gpurun_out/c_pytest.log
timeout 400 python scripts/bench_gemm_kinds.py > gpurun_out/c_gemm_kinds.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c_bench_tds.json 2> gpurun_out/c_bench_tds.err
timeout 300 python bench.py --workload conv_glu_asg --steps 5 --warmup 8 --no-cpu > gpurun_out/c_bench_convglu.json 2> gpurun_out/c_bench_convglu.err
timeout 300 python bench.py --workload streaming_tds_ctc --steps 5 --warmup 5 --no-cpu > gpurun_out/c_bench_streaming.json 2> gpurun_out/c_bench_streaming.err
timeout 300 ncu --set full --clock-control none -k regex:conv_umma_fwd -c 3 -o /tmp/c_prof_conv python scripts/prof_conv.py > gpurun_out/c_ncu_conv.log 2>&1
ncu -i /tmp/c_prof_conv.ncu-rep --page raw --csv > gpurun_out/c_prof_conv_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:gemm_umma -c 3 -o /tmp/c_prof_gemm_bf16 python scripts/prof_gemm.py bf16 > gpurun_out/c_ncu_bf16.log 2>&1
ncu -i /tmp/c_prof_gemm_bf16.ncu-rep --page raw --csv > gpurun_out/c_prof_gemm_bf16_raw.csv 2>/dev/null
du -sh gpurun_out
