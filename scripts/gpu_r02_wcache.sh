#!/bin/bash
mkdir -p gpurun_out
for cfg in "c4 1024 heun 2000" "c3 1024 euler 2000" "c3 1024 rk4 1000" "c4 512 heun 2000"; do
  set -- $cfg
  for flags in "" "--no-wcache"; do
    timeout 300 python bench_configs.py --config $1 --size $2 --solver $3 --steps $4 --check --cpu-steps 5 $flags >> gpurun_out/r02_wcache.jsonl 2>> gpurun_out/r02_wcache.err
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_wcache.jsonl'):
    d=json.loads(l); print(d['config'], d['n'], d['solver'], 'wcache' if d['w_stationary'] else 'stream', round(d['ms_per_step']*1e3,3), 'us', d.get('max_abs_diff_vs_oracle_obs0'))
PY
timeout 600 python -m pytest tests/test_gpu_network.py tests/test_gpu_fullsize.py -x -q -m gpu -p no:cacheprovider > gpurun_out/r02_wcache_tests.txt 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r02_wcache_tests.txt
