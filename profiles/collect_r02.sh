#!/bin/bash
# Round-2 captures on one B200 (run through gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash profiles/collect_r02.sh r02q'
# 1. ncu launch list of a short bench run at C3, 2. `ncu --set full` of one iteration's GEMM / sweep launches,
# 3. `ncu --set full` of the NNLS kernels of a C3-shaped dual regression, 4. the per-GPU shard of C3 at 8 GPUs
# (7427 x 110000) on one GPU: plain timed run, then its launch list -- event-timed iteration against the sum of
# the kernels' own durations = what the ~26 kernel boundaries of an iteration cost.  Outputs land in gpurun_out/.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
B="python bench.py --no-e2e --no-cpu --no-side --steps 2 --warmup 3"
timeout 300 $B > /dev/null 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:nfb:: \
    -c 400 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_split_kernel|cd_sweep_tile' \
    --launch-skip 12 -c 6 -o $out/${tag}_full -f $B > $out/${tag}_ncu_full.log 2>&1
timeout 120 python tests/gpu_dr_probe.py C3 1 > /dev/null 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'nnls_greedy_kernel|nnls_smem_kernel' -c 4 \
    -o $out/${tag}_nnls_full -f python tests/gpu_dr_probe.py C3 1 > $out/${tag}_ncu_nnls.log 2>&1
S="python bench.py --shape 7427,110000,200 --no-e2e --no-cpu --no-side --steps 20 --warmup 5"
NFB_PHASE_TIMING=1 timeout 300 $S > $out/${tag}_shard8_bench.json 2> $out/${tag}_shard8_bench.err
timeout 300 $S > $out/${tag}_shard8_bench_plain.json 2> /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:nfb:: \
    -c 400 --csv --log-file $out/${tag}_shard8_launches.csv \
    python bench.py --shape 7427,110000,200 --no-e2e --no-cpu --no-side --steps 2 --warmup 3 > $out/${tag}_shard8_ncu.log 2>&1
ls -la $out | tail -12
