#!/bin/bash
# One GPU-box visit for the committed evidence of a round: bench (all configs, both arms), ncu launch list and full
# captures of the top kernels.  Usage (through gpurun): bash tools/gpu_profile_round.sh <tag>
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 900 python bench.py --steps 20 --warmup 5 2>$O/${TAG}_bench.err | tail -1 > $O/bench_${TAG}.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>>$O/${TAG}_bench.err | tail -1 > $O/bench_${TAG}_reference_arm.json
timeout 600 python bench.py --config cfg2 --steps 10 --warmup 3 2>>$O/${TAG}_bench.err | tail -1 > $O/bench_${TAG}_cfg2.json
timeout 600 python bench.py --config cfg4 --steps 10 --warmup 3 2>>$O/${TAG}_bench.err | tail -1 > $O/bench_${TAG}_cfg4.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-torch-gpu > $O/${TAG}_ncu_bench.log 2>&1
for k in attention_h3s_kernel gemm_tc_persist_kernel sinkhorn_cl_kernel mvba_kernel ba_init_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 2 -f -o $O/${TAG}_$k \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-torch-gpu > $O/${TAG}_ncu_$k.log 2>&1
done
python - <<P
import json
for f in ('bench_${TAG}.json', 'bench_${TAG}_cfg2.json', 'bench_${TAG}_cfg4.json'):
    try:
        d=json.load(open('$O/' + f)); print(f, {k:d[k] for k in ('value','ms_per_step','e2e','clocks','roofline','roofline_sinkhorn','stage_ms_per_step','cpu_baseline','pose_auc_parity','pose_auc_5_10_20') if k in d})
    except Exception as e:
        print(f, 'ERR', e)
print(open('$O/bench_${TAG}_reference_arm.json').read()[:700])
P
tail -c 1500 $O/${TAG}_bench.err
