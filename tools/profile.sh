#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, 1 GPU).  $1 = tag (e.g. r01c)
TAG=${1:-r01}
mkdir -p gpurun_out
# (1) launch list of eager 720p frames: skip the warm-up frames
ncu --metrics gpu__time_duration.sum --clock-control none -s 2700 -c 1100 --csv \
    --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --frames 2 --no-cpu-baseline --no-graph > gpurun_out/${TAG}_ncu_bench.log 2>&1
# (2) full captures of the top kernels
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 21 -c 2 \
    -o gpurun_out/${TAG}_prof_gemm_tc_conv -f python tools/kbench.py --only conv --iters 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 5 -c 2 \
    -o gpurun_out/${TAG}_prof_gemm_tc_linear -f python tools/kbench.py --only gemm --iters 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:msda_kernel -s 3 -c 1 \
    -o gpurun_out/${TAG}_prof_msda -f python tools/kbench.py --only msda --iters 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 11 -c 1 \
    -o gpurun_out/${TAG}_prof_attn -f python tools/kbench.py --only attn --iters 1 > /dev/null 2>&1
ls -la gpurun_out/ | grep ${TAG}
