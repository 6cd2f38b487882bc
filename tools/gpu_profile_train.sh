#!/bin/bash
# One GPU-box visit for the evidence of the TRAINING step (bench --config cfg5): ncu launch list of one step and full
# captures of the attention-backward and Sinkhorn-training kernels.  Usage (through gpurun): bash tools/gpu_profile_train.sh <tag>
TAG=${1:-r02_train}
O=gpurun_out
mkdir -p $O
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --config cfg5 --steps 1 --warmup 3 > $O/${TAG}_ncu_bench.log 2>&1
for k in attn_bwd_dq_mma_kernel attn_bwd_dkv_mma_kernel sinkhorn_train_bwd_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 30 -c 1 -f -o $O/${TAG}_$k \
      python bench.py --config cfg5 --steps 1 --warmup 3 > $O/${TAG}_ncu_$k.log 2>&1
done
ls -la $O | grep ${TAG}
