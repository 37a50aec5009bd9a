#!/bin/bash
# First GPU call of the next round: run everything that was written after round 1's GPU budget was spent
# (tests gated behind DSB200_RUN_UNVERIFIED=1) and measure it against the default paths.
#   bash tools/verify_unverified.sh          one B200:  gated single-GPU tests, GEMM loader 3 timing, e2e with / without pinned_mirror
#   bash tools/verify_unverified.sh multi N  N B200s:   gated multi-GPU tests, bench with NCCL vs the peer-memory exchange kernels
mkdir -p gpurun_out
if [ "$1" = "multi" ]; then
    N=${2:-2}
    DSB200_RUN_UNVERIFIED=1 timeout 900 python -m pytest tests/test_multi_gpu.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/unverified_multi.log
    for p2p in 0 1; do
        timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + p2p)) \
            bench.py --gpus $N --steps 50 --warmup 5 --cpu-steps 0 --p2p $p2p > gpurun_out/bench_mp${N}_p2p${p2p}.json 2> gpurun_out/bench_mp${N}_p2p${p2p}.err
        python - <<PY
import json
t = open("gpurun_out/bench_mp${N}_p2p${p2p}.json").read(); i = t.find('{"metric')
d = json.loads(t[i:].splitlines()[0]) if i >= 0 else {}
print("mp${N} p2p=${p2p}:", d.get("value"), d.get("ms_per_step"), {k: round(v["ms_per_call"] * 1e3, 1) for k, v in d.get("kernels", {}).items() if k in ("all_gather", "reduce_scatter")})
PY
    done
    exit 0
fi
DSB200_RUN_UNVERIFIED=1 timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/unverified_single.log
timeout 120 python tools/gemm_debug_matrix.py 2 3 2>&1 | grep -E "loader|^ +[23] +0 " | tee gpurun_out/gemm_loader3.log
timeout 120 python tools/fused_output_bench.py 2>&1 | tee gpurun_out/fused_output_bench.log
timeout 300 python bench.py --steps 100 --warmup 10 --cpu-steps 0 --fuse-output 1 > gpurun_out/bench_fuse1.json 2> gpurun_out/bench_fuse1.err
python -c "import json; d=json.load(open('gpurun_out/bench_fuse1.json')); print('fuse_output_gemm=1: value', d['value'], 'ms', d['ms_per_step'])"
for pm in 0 1; do
    timeout 300 python bench.py --steps 100 --warmup 10 --cpu-steps 0 --pinned-mirror $pm > gpurun_out/bench_pm${pm}.json 2> gpurun_out/bench_pm${pm}.err
    python -c "import json; d=json.load(open('gpurun_out/bench_pm${pm}.json')); print('pinned_mirror=${pm}: value', d['value'], 'e2e', d['e2e'])"
done
