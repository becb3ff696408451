#!/bin/bash
# round-2 GPU session H: ncu of the new ASG kernels (launch list + full set + source page of the chains kernel)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/h_asg_launches.csv python scripts/prof_asg.py asg 3 > gpurun_out/h_ncu1.log 2>&1
tail -12 gpurun_out/h_asg_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:asg_ -s 10 -c 5 -o /tmp/h_asg python scripts/prof_asg.py asg 3 > gpurun_out/h_ncu2.log 2>&1
ncu -i /tmp/h_asg.ncu-rep --page raw --csv > gpurun_out/h_asg_raw.csv 2>/dev/null
ncu -i /tmp/h_asg.ncu-rep --page source --csv -k regex:asg_chains > gpurun_out/h_asg_chains_source.csv 2>/dev/null
python scripts/ncu_summary.py /tmp/h_asg.ncu-rep > gpurun_out/h_asg_summary.txt 2>&1
python scripts/ncu_top_stalls.py gpurun_out/h_asg_chains_source.csv 40 > gpurun_out/h_asg_chains_top.txt 2>&1
head -50 gpurun_out/h_asg_chains_top.txt
du -sh gpurun_out
