#!/bin/bash
# round-2 GPU session E: in-step GEMM diagnosis (per-launch times, ncu of in-step launches), conv_glu after the kernel tuning
mkdir -p gpurun_out && rm -f gpurun_out/arch_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/e_pytest.log
tail -5 gpurun_out/e_pytest.log
for p in f32 tf32 bf16; do
  timeout 300 python bench.py --precision $p --steps 10 --warmup 3 --no-extras --no-cpu > gpurun_out/e_bench_tds_$p.json 2> gpurun_out/e_bench_tds_$p.err
done
timeout 300 python bench.py --workload conv_glu_asg --steps 5 --warmup 8 --no-cpu > gpurun_out/e_bench_convglu.json 2> gpurun_out/e_bench_convglu.err
timeout 300 python bench.py --workload streaming_tds_ctc --steps 5 --warmup 5 --no-cpu > gpurun_out/e_bench_streaming.json 2> gpurun_out/e_bench_streaming.err
# in-step GEMM launches of the 4th step (3 warm-up steps x 69 GEMMs skipped): forward + backward of the last stage
timeout 600 ncu --set full --clock-control none -k regex:gemm_umma -s 207 -c 69 -o /tmp/e_prof_instep python bench.py --precision tf32 --steps 1 --warmup 3 --no-extras --no-cpu > gpurun_out/e_ncu_instep.log 2>&1
ncu -i /tmp/e_prof_instep.ncu-rep --page raw --csv > gpurun_out/e_prof_instep_raw.csv 2>/dev/null
du -sh gpurun_out
