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
d=json.load(open('gpurun_out/r02_bench_n$N.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('d2h_ceiling_gbs'), d['e2e'].get('frac_of_ceiling'), 'red', d['e2e_reduced']['value'], 'jit', d['jit_s'], 'alloc', d['alloc_s'])
print('strong', d.get('strong_scaling'))
for k,v in d.get('configs',{}).items(): print(k, v.get('value'), v.get('us_per_solver_step', v.get('ms_per_solver_step')), v.get('roofline',{}).get('frac'), v.get('error'))
PY
