#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests2.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests2.log
tail -8 gpurun_out/r2_tests2.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_full.log 2> gpurun_out/r2_bench_full.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench_full.err
python - <<'PY'
import json
try:
    l = json.loads(open('gpurun_out/r2_bench_full.log').read().strip().splitlines()[-1])
    print('value', l['value'], 'ms', l['ms_per_step'], 'e2e', l['e2e'], 'frac', l['roofline']['frac'])
    print('predict', {k: l['predict'][k] for k in ('value','e2e','ms_per_step')}, l['predict']['roofline']['frac'], l['predict'].get('cpu_baseline'))
    print('cpu', l.get('cpu_baseline'))
    for k, v in l.get('configs', {}).items():
        if 'error' in v: print(k, v); continue
        print(k, 'build', v['build'], 'e2e', v['e2e'], 'frac', v['roofline']['frac'], 'pred', v['predict']['value'], v['predict']['roofline']['frac'], 'wall', v['wall_s'])
        print('   cpu', v['cpu_baseline'], v['predict'].get('cpu_baseline'))
        print('   stats', v['stats_per_step'])
except Exception as e:
    print('parse failed', e)
PY
