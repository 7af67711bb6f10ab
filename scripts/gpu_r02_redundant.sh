#!/bin/bash
# exchange-form A/B on N GPUs + the multi-GPU parity test
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multigpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/r02_redundant_tests_n$N.txt 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r02_redundant_tests_n$N.txt
for cfg in "c3 2048 euler 2000" "c3 1024 euler 2000" "c3 4096 rk4 500" "c4 512 heun 2000"; do
  set -- $cfg
  for flags in "" "--no-redundant"; do
    timeout 200 $TR --master-port 29551 bench_configs.py --config $1 --size $2 --solver $3 --steps $4 --check --cpu-steps 5 $flags >> gpurun_out/r02_redundant_n$N.jsonl 2>> gpurun_out/r02_redundant_n$N.err
  done
done
python - <<PY
import json
for l in open('gpurun_out/r02_redundant_n$N.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print(d['config'], d['n'], d['solver'], d['n_gpus'], d['exchange'], round(d['ms_per_step']*1e3,3), 'us', d.get('max_abs_diff_vs_oracle_obs0'))
PY
