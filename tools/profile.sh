#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, 1 GPU).  $1 = tag (e.g. r01q).  Summaries:
#   python tools/ncu_summary.py launches gpurun_out/${TAG}_launches.csv profiles/${TAG}_launches_graph_b20.json "<note>"
#   python tools/gemm_traffic.py gpurun_out/${TAG}_gemm_traffic.csv profiles/${TAG}_gemm_traffic.json 20
#   python tools/ncu_summary.py full gpurun_out/${TAG}_prof_<k>.ncu-rep profiles/${TAG}_ncu_<k>.json
TAG=${1:-r01}
mkdir -p gpurun_out
# (1) launch list of ONE replay of the 20-frame 720p graph (exactly the kernels bench.py times)
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${TAG}_launches.csv python tools/ncu_frame.py 20 > gpurun_out/${TAG}_ncu_frame.log 2>&1
# (2) DRAM traffic of every GEMM launch of that replay (bench.py roofline.traffic)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --profile-from-start off -k regex:gemm_tc_kernel --csv --log-file gpurun_out/${TAG}_gemm_traffic.csv \
    python tools/ncu_frame.py 20 > gpurun_out/${TAG}_traffic.log 2>&1
# (3) full captures of the top kernels (8-frame shapes: quick to replay ~40 times)
full() {  # name, kernel regex, skip, command...
    local name=$1 rx=$2 skip=$3; shift 3
    ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 \
        -o gpurun_out/${TAG}_prof_$name -f "$@" > gpurun_out/${TAG}_prof_$name.log 2>&1
}
full gemm_tc_conv gemm_tc_kernel 3 python tools/kbench.py --iters 1 --batch 8 --only conv --match '256->256@184x320'
full msda msda_group_kernel 3 python tools/kbench.py --iters 1 --batch 8 --only msda
# kernels inside the frame graph: profile between cudaProfilerStart/Stop of tools/ncu_frame.py
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_mma_kernel -s 8 -c 1 \
    -o gpurun_out/${TAG}_prof_attn -f python tools/ncu_frame.py 8 > gpurun_out/${TAG}_prof_attn.log 2>&1
ls -la gpurun_out/ | grep ${TAG}
# (4) Swin-B backbone (bench.py --backbone swin_b): launch list of one 8-frame replay, window-attention and
#     overlap-kernel captures (profiles/r01s_*, r01u_*):
#   python tools/ncu_summary.py launches gpurun_out/${TAG}_swin_launches.csv profiles/${TAG}_swin_b_launches_graph_b8.json "<note>"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${TAG}_swin_launches.csv python tools/ncu_frame.py 8 swin_b > gpurun_out/${TAG}_swin_ncu_frame.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:window_attention_mma -s 3 -c 1 \
    -o gpurun_out/${TAG}_prof_winatt -f python tools/swin_time.py 4 > gpurun_out/${TAG}_prof_winatt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:overlap_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof_overlap -f \
    python -c "import bench, torch; bench.relset_bench(torch.device('cuda'), False)" > gpurun_out/${TAG}_prof_overlap.log 2>&1
ls -la gpurun_out/ | grep ${TAG}
