#!/bin/bash
# Round profile on the GPU box (one B200):  bash tools/profile_round.sh <tag> [full]
#   1. launch list of bench.py (gpu__time_duration.sum, no clock control)      -> gpurun_out/launches_<tag>.csv
#   2. ncu --set full of the three output-layer GEMMs alone (tools/gemm_once)  -> gpurun_out/gemm_<tag>.ncu-rep
#   3. (only with "full") ncu --set full of the hand-written kernels of one training step inside bench.py
#                                                                               -> gpurun_out/prof_<tag>.ncu-rep
# gpurun brings back at most 64 MiB: keep --import-source to the small captures.  Numbers printed by bench.py under
# ncu are never bench values; summaries are made with tools/summarize_profiles.py.
tag=${1:-r1c}
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 99 -c 300 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 4 --warmup 3 --cpu-steps 0 > gpurun_out/ncu_launches_${tag}.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip 3 --launch-count 3 -f -o gpurun_out/gemm_${tag} \
    python tools/gemm_once.py 2 > gpurun_out/ncu_gemm_${tag}.log 2>&1
if [ "$2" = "full" ]; then
    timeout 400 ncu --set full --clock-control none -k regex:'gemm_tc|output_row|sparse_|update_biases|transpose_scatter' --launch-skip 40 --launch-count 12 \
        -f -o gpurun_out/prof_${tag} python bench.py --steps 4 --warmup 3 --cpu-steps 0 > gpurun_out/ncu_full_${tag}.log 2>&1
fi
ls -la gpurun_out/*${tag}*
