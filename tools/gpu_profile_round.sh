#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list and full captures of the top kernels.
# Usage (through gpurun): bash tools/gpu_profile_round.sh <tag>
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu --timeout 300 --tb=line 2>&1 | tail -6 > $O/${TAG}_pytest.txt
cat $O/${TAG}_pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 2>$O/${TAG}_bench.err | tail -1 > $O/bench_${TAG}.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>>$O/${TAG}_bench.err | tail -1 > $O/bench_${TAG}_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
for k in attention_tc_kernel gemm_tc_persist_kernel sinkhorn_exp_kernel mvba_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 2 -f -o $O/${TAG}_$k \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu_$k.log 2>&1
done
python - <<P
import json
d=json.load(open('$O/bench_${TAG}.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','clocks','roofline','roofline_sinkhorn','stage_ms_per_step','cpu_baseline','tf32_single_pass') if k in d})
print(open('$O/bench_${TAG}_reference_arm.json').read()[:600])
P
