#!/bin/bash
# round-2 GPU session B: validate the persistent GEMM, 3xTF32 conv, fused LN; timings per precision; ncu of the GEMM
mkdir -p gpurun_out && rm -f gpurun_out/arch_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b_pytest.log
tail -5 gpurun_out/b_pytest.log
timeout 300 python scripts/bench_gemm_kinds.py > gpurun_out/b_gemm_kinds.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/b_bench_tds.json 2> gpurun_out/b_bench_tds.err
timeout 300 python bench.py --workload conv_glu_asg --steps 5 --no-cpu > gpurun_out/b_bench_convglu.json 2> gpurun_out/b_bench_convglu.err
timeout 300 python bench.py --workload streaming_tds_ctc --steps 5 --no-cpu > gpurun_out/b_bench_streaming.json 2> gpurun_out/b_bench_streaming.err
for k in tf32 bf16 f32x3; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_umma -c 3 -o /tmp/b_prof_gemm_$k python scripts/prof_gemm.py $k > gpurun_out/b_ncu_$k.log 2>&1
  ncu -i /tmp/b_prof_gemm_$k.ncu-rep --page raw --csv > gpurun_out/b_prof_gemm_${k}_raw.csv 2>/dev/null
  ncu -i /tmp/b_prof_gemm_$k.ncu-rep --page details --csv > gpurun_out/b_prof_gemm_${k}_details.csv 2>/dev/null
done
cp /tmp/b_prof_gemm_bf16.ncu-rep gpurun_out/ 2>/dev/null   # one full report (source page) comes home; the others as CSV
du -sh gpurun_out
ls -la gpurun_out | grep " b_"
