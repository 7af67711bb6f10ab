#!/bin/bash
# Round 2, 8-GPU call: multi-GPU parity at world 2 / 4 / 8, the full bench line at N = 8, the D2H ceiling probe.
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1; nproc >> gpurun_out/r02_topo_n$N.txt; numactl -H >> gpurun_out/r02_topo_n$N.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multigpu.py -x -q -m gpu -p no:cacheprovider -rs > gpurun_out/r02_multigpu_tests_n$N.txt 2>&1
echo "multigpu tests rc=$?"; tail -6 gpurun_out/r02_multigpu_tests_n$N.txt
timeout 300 $TR --master-port 29521 scripts/d2h_probe.py > gpurun_out/r02_d2h_probe_n$N.json 2> gpurun_out/r02_d2h_probe_n$N.err
echo "d2h rc=$?"; cat gpurun_out/r02_d2h_probe_n$N.json
timeout 1200 $TR --master-port 29522 bench.py --gpus $N > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/r02_bench_n$N.json') if l.startswith('{')][-1]
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('d2h_ceiling_gbs'), d['e2e'].get('frac_of_ceiling'), 'red', d['e2e_reduced']['value'], 'jit', d['jit_s'], 'alloc', d['alloc_s'])
print('strong', d.get('strong_scaling'))
for k,v in d.get('configs',{}).items(): print(k, v.get('value'), v.get('us_per_solver_step', v.get('ms_per_solver_step')), v.get('roofline',{}).get('frac'), v.get('error'))
PY
if [ "$N" = "8" ]; then
  # small networks across 8 GPUs (latency regime): N = 2048 / 1024 dense WC, QIF-SFA 2 x 1024
  for cfg in "c3 2048 euler 2000" "c3 1024 euler 2000" "c3 4096 rk4 500" "c3 16384 euler 200"; do
    set -- $cfg
    timeout 300 $TR --master-port 29523 bench_configs.py --config $1 --size $2 --solver $3 --steps $4 --check --cpu-steps 5 >> gpurun_out/r02_cfg_n8.jsonl 2>> gpurun_out/r02_cfg_n8.err
  done
  cut -c1-260 gpurun_out/r02_cfg_n8.jsonl
fi
