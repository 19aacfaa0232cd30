#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, 1 GPU).  $1 = tag (e.g. r01e)
TAG=${1:-r01}
mkdir -p gpurun_out
# (1) launch list of ONE replay of the batch-8 720p frame graph (the kernels bench.py times)
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${TAG}_launches.csv python tools/ncu_frame.py 8 > gpurun_out/${TAG}_ncu_frame.log 2>&1
# (2) full captures of the top kernels at the batch-8 shapes
full() {  # name, kernel regex, skip, kbench args...
    local name=$1 rx=$2 skip=$3; shift 3
    ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 \
        -o gpurun_out/${TAG}_prof_$name -f python tools/kbench.py --iters 1 --batch 8 "$@" > gpurun_out/${TAG}_prof_$name.log 2>&1
}
full gemm_tc_conv gemm_tc_kernel 3 --only conv --match '256->256@184x320'
full gemm_tc_linear gemm_tc_kernel 3 --only gemm --match 'linear_154560x1024x256'
full msda msda_kernel 3 --only msda
full attn attn_kernel 3 --only attn --match '100x14720'
full pan pan_pixel_kernel 1 --only pan
ls -la gpurun_out/ | grep ${TAG}
