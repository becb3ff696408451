#!/bin/bash
# round-2: sliced (halo) FAC gradient for long targets: parity, memcheck, sweep timing
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_criterion.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/halo_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/halo_pytest.log
tail -8 gpurun_out/halo_pytest.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_criterion.py -m gpu -q --tb=line -p no:cacheprovider -k "test_asg_parity and not baseline" > gpurun_out/halo_memcheck.log 2>&1; tail -3 gpurun_out/halo_memcheck.log
timeout 600 python bench.py --workload asg_sweep --steps 10 --warmup 3 --no-cpu > gpurun_out/halo_sweep.json 2> gpurun_out/halo_sweep.err; tail -2 gpurun_out/halo_sweep.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/halo_sweep.json').read().strip().splitlines()[-1])
for p in d['points']:
    if p['T']>=1500: print(p['T'],p['B'],p['L'],round(p['ms_per_batch'],3),'step ns',round(p['dependent_step_ns'] or 0),{k:float(f'{v:.1e}') for k,v in p['oracle_parity_rel'].items()})
PY
timeout 400 python scripts/asg_parity_seeds.py 4
timeout 100 python scripts/asg_roles.py | tr -d "\n" | cut -c1-420
