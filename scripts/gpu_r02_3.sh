#!/bin/bash
mkdir -p gpurun_out
python scripts/pinned_probe.py > gpurun_out/r02_pinned_probe.txt 2>&1; cat gpurun_out/r02_pinned_probe.txt
for cfg in "c4 1024 heun 2000" "c3 2048 euler 2000" "c3 1024 euler 2000" "c3 8192 euler 500" "c3 8192 rk4 200" "c4 2048 heun 1000" "c5 8192 euler 200"; do
  set -- $cfg
  for flags in "" "--no-split" "--no-deep" "--no-split --no-deep"; do
    timeout 300 python bench_configs.py --config $1 --size $2 --solver $3 --steps $4 --check --cpu-steps 5 $flags >> gpurun_out/r02_ab.jsonl 2>> gpurun_out/r02_ab.err
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_ab.jsonl'):
    d=json.loads(l); print(d['config'], d['n'], d['solver'], 'split' if d['split_rows'] else 'nosplit', 'deep' if d['dot_deep'] else 'nodeep', round(d['ms_per_step']*1e3,3), 'us', round(d['roofline']['frac'],3), d.get('max_abs_diff_vs_oracle_obs0'))
PY
timeout 600 python bench.py --configs c3,c3_rk4,c4,c5,public_api --no-cpu-baseline > gpurun_out/r02_bench_part.json 2> gpurun_out/r02_bench_part.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_part.json'))
for k,v in d['configs'].items(): print(k, v.get('us_per_solver_step'), v.get('roofline',{}).get('frac'), v.get('clocks'), v.get('error'))
print(d['public_api']); print('alloc_s', d['alloc_s'], 'jit_s', d['jit_s'])
PY
timeout 600 python scripts/tune_inst.py > gpurun_out/r02_tune_inst.jsonl 2>&1; cat gpurun_out/r02_tune_inst.jsonl
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_named_sizes.py tests/test_gpu_fullsize.py -x -q -m gpu -p no:cacheprovider > gpurun_out/r02_regress.txt 2>&1
echo "regress rc=$?"; tail -3 gpurun_out/r02_regress.txt
