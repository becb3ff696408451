#!/bin/bash
# round-2 multi-GPU session: N = $1 ranks on one box (torchrun, NCCL): TDS+CTC step, conv_glu+ASG step (836 MB arena), ASG
N=${1:-2}
WL=${2:-"tds_ctc conv_glu_asg asg_sweep"}
mkdir -p gpurun_out
for wl in $WL; do
  extra="--no-cpu"; [ "$wl" = "tds_ctc" ] && extra="--no-cpu --no-extras"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $wl --steps 10 --warmup 3 $extra > gpurun_out/multi_n${N}_$wl.json 2> gpurun_out/multi_n${N}_$wl.err
  tail -c 400 gpurun_out/multi_n${N}_$wl.json | head -c 400; echo; tail -2 gpurun_out/multi_n${N}_$wl.err
done
