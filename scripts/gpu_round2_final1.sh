#!/bin/bash
# round-2 final single-GPU session: full GPU suite, every bench workload, ncu launch lists and full-set captures
mkdir -p gpurun_out && rm -f gpurun_out/arch_parity.jsonl
( time timeout 1500 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider ) > gpurun_out/fin_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/fin_pytest.log
tail -6 gpurun_out/fin_pytest.log
timeout 600 python bench.py > gpurun_out/fin_bench_tds_ctc.json 2> gpurun_out/fin_bench_tds_ctc.err
for wl in conv_glu_asg streaming_tds_ctc tds_asg asg asg_sweep; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu > gpurun_out/fin_bench_$wl.json 2> gpurun_out/fin_bench_$wl.err
done
tail -c 300 gpurun_out/fin_bench_*.err
# ncu: launch list of the default command (2 steps), launch list + full set of the criterion kernels
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/fin_tds_ctc_launches.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu > gpurun_out/fin_ncu_step.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"asg_|ctc_" -s 12 -c 6 -o /tmp/fin_asg python scripts/prof_asg.py asg 3 > gpurun_out/fin_ncu_asg.log 2>&1
python scripts/ncu_summary.py /tmp/fin_asg.ncu-rep > gpurun_out/fin_asg_summary.txt 2>&1
du -sh gpurun_out
