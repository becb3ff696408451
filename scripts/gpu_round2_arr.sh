#!/bin/bash
# round-2: conv1d arrange / unarrange with incremental indexing: parity + conv_glu step time
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_convglu.py tests/test_gpu_archs.py -m gpu -q --tb=short -p no:cacheprovider -x -k "conv or glu" > gpurun_out/arr_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/arr_pytest.log
tail -5 gpurun_out/arr_pytest.log
timeout 200 python bench.py --workload conv_glu_asg --steps 10 --warmup 3 --no-cpu > gpurun_out/arr_bench.json 2> gpurun_out/arr_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/arr_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], [(k,round(v['ms'],3)) for k,v in list(d['step_breakdown']['kernels'].items())[:8]])
PY
