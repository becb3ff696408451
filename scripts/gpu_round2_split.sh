#!/bin/bash
# round-2: f32x3 GEMM with 8 vs 12 split warps (W2L_F32X3_SPLIT_THREADS): accuracy tests of the 12-warp variant, step time of both
mkdir -p gpurun_out
W2L_F32X3_SPLIT_THREADS=384 timeout 200 python -m pytest tests/test_gpu_gemm.py -m gpu -q --tb=line -p no:cacheprovider -x > gpurun_out/split_pytest.log 2>&1; tail -2 gpurun_out/split_pytest.log
for st in 256 384; do
W2L_F32X3_SPLIT_THREADS=$st timeout 200 python bench.py --precision f32 --steps 10 --warmup 3 --no-extras --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('split threads $st: step', round(d['ms_per_step'],3), 'gemm', round(d['roofline']['gemm_ms_per_step'],3), 'loss', d['final_loss_sum'])"
done
