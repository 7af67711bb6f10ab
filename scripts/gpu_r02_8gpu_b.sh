#!/bin/bash
# Round 2, last 8-GPU call: multi-GPU parity incl. the two-level vsum reduction at world 2 / 4 / 8; NVLink byte counters around
# one row-sharded C3 run (ncu cannot wrap a multi-rank command: nvidia-smi's per-link counters are the evidence instead).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multigpu.py -x -q -m gpu -p no:cacheprovider -rs > gpurun_out/r02_multigpu_tests_n8_final.txt 2>&1
echo "multigpu tests rc=$?"; tail -4 gpurun_out/r02_multigpu_tests_n8_final.txt
nvidia-smi nvlink -gt d -i 0 > gpurun_out/r02_nvlink_before.txt 2>&1
timeout 300 $TR --master-port 29541 bench_configs.py --config c3 --size 8192 --solver rk4 --steps 2000 --cpu-steps 0 > gpurun_out/r02_nvlink_run.json 2> gpurun_out/r02_nvlink_run.err
nvidia-smi nvlink -gt d -i 0 > gpurun_out/r02_nvlink_after.txt 2>&1
tail -1 gpurun_out/r02_nvlink_run.json | cut -c1-250
python - <<'PY'
import re
def tot(p):
    tx=rx=0
    for l in open(p):
        m=re.search(r'Data (Tx|Rx): (\d+) KiB', l)
        if m:
            if m.group(1)=='Tx': tx+=int(m.group(2))
            else: rx+=int(m.group(2))
    return tx, rx
b, a = tot('gpurun_out/r02_nvlink_before.txt'), tot('gpurun_out/r02_nvlink_after.txt')
print('GPU0 NVLink data KiB before', b, 'after', a, 'delta tx/rx KiB', a[0]-b[0], a[1]-b[1])
PY
timeout 600 $TR --master-port 29542 bench.py --gpus 8 --configs c5a,c4 --no-cpu-baseline > gpurun_out/r02_bench_n8_c5a.json 2> gpurun_out/r02_bench_n8_c5a.err
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r02_bench_n8_c5a.json') if l.startswith('{')][-1]
for k,v in d['configs'].items(): print(k, v.get('value'), v.get('us_per_solver_step'), v.get('error'))
PY
