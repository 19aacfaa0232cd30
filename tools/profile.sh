#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, 1 GPU).  $1 = tag (e.g. r01)
TAG=${1:-r01}
mkdir -p gpurun_out
# (1) launch list of one eager frame: skip 2 warm-up frames (~640 launches each)
ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 660 --csv \
    --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --frames 2 --no-cpu-baseline --no-graph > gpurun_out/${TAG}_ncu_bench.log 2>&1
# (2) full captures of the top kernels
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 40 -c 3 \
    -o gpurun_out/${TAG}_prof_gemm -f python tools/kbench.py --only conv --iters 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:msda_kernel -s 3 -c 2 \
    -o gpurun_out/${TAG}_prof_msda -f python tools/kbench.py --only msda --iters 1 > /dev/null 2>&1
ls -la gpurun_out/
