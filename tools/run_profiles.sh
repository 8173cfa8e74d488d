#!/bin/bash
# GPU box: the two ncu passes of /opt/skills/guides/B200_PROFILING.md on the current build (each after the same command
# exited 0 without ncu). Outputs under gpurun_out/: launches_<tag>.csv (launch list of one captured UNet step) and
# prof_shapes_<tag>.ncu-rep (--set full of representative GEMM / conv / attention launches).
TAG=${1:-r2}
N=${2:-380}   # kernel launches per captured step (bench.py: gpu_launches_per_step)
K='regex:tf_|gn_|ln_|conv3x3_|gemv_|upsample|cfg_|pad_tok|timestep|add_int'
python bench.py --quick --steps 2 --warmup 3 > gpurun_out/plain_bench.log 2>&1 && \
ncu -k "$K" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -s $N -c $N --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_launch_${TAG}.log 2>&1
python tools/profile_shapes.py > gpurun_out/plain_shapes.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:tf_gemm_kernel|tf_attention' -o gpurun_out/prof_shapes_${TAG} \
    python tools/profile_shapes.py > gpurun_out/ncu_shapes_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_launch_${TAG}.log gpurun_out/ncu_shapes_${TAG}.log
ls -la gpurun_out/
