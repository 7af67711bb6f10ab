#!/bin/bash
# Round 2, call 5 (1 GPU): tcgen05 stage-granularity A/B + parity, then ncu: launch list + full captures of the three kernels.
mkdir -p gpurun_out
for ck in 32 16; do
  PRB_TC_CK=$ck timeout 300 python bench_configs.py --config c3 --size 8192 --batch 1024 --precision float32 --sweep k,tau --steps 20 --cpu-steps 0 >> gpurun_out/r02_tc_ck.jsonl 2>> gpurun_out/r02_tc_ck.err
  PRB_TC_CK=$ck timeout 300 python bench_configs.py --config c3 --size 8192 --batch 512 --precision float32 --sweep k,tau --steps 20 --cpu-steps 0 >> gpurun_out/r02_tc_ck.jsonl 2>> gpurun_out/r02_tc_ck.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_tc_ck.jsonl'):
    d=json.loads(l); print('tc', d['batch'], d['tc'], round(d['ms_per_step'],4), 'ms', d['roofline']['achieved'], d['roofline']['frac'])
PY
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_named_sizes.py -x -q -m gpu -p no:cacheprovider -k "tensor_core or tcgen05 or sweep" > gpurun_out/r02_tc_tests.txt 2>&1
echo "tc tests rc=$?"; tail -3 gpurun_out/r02_tc_tests.txt
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --configs c3,c3_sweep_f32"
$CMD > gpurun_out/r02_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:^prb_integrate$ -s 40 -c 1 -o gpurun_out/r02_c2_prb_integrate $CMD > gpurun_out/r02_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:prb_integrate_net -s 3 -c 1 -o gpurun_out/r02_c3_prb_integrate_net $CMD > gpurun_out/r02_ncu_c3.log 2>&1
echo "ncu c3 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:prb_integrate_net -s 11 -c 1 -o gpurun_out/r02_tc_prb_integrate_net $CMD > gpurun_out/r02_ncu_tc.log 2>&1
echo "ncu tc rc=$?"
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/r02_ncu_plain.log | cut -c1-300
