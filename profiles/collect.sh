#!/bin/bash
# Round-end measurement on one B200 (run through gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash profiles/collect.sh r01b'
# 1. GPU tests, 2. the bench line the driver records, 3. the ncu launch list of a short bench run,
# 4. one `ncu --set full` capture of the dominant kernels, 5. one of the NNLS kernel.  Outputs land in gpurun_out/
# (scratch);
# profiles/summarise.py turns them into the tracked summaries.
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -2 $out/${tag}_pytest.log
timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err || tail -5 $out/${tag}_bench.err
tail -c 600 $out/${tag}_bench.json
timeout 300 python bench.py --no-e2e --no-cpu --steps 2 --warmup 3 > /dev/null 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:nfb:: \
    -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --no-e2e --no-cpu --steps 2 --warmup 3 > $out/${tag}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_split_kernel|cd_sweep_tile' \
    --launch-skip 12 -c 6 -o $out/${tag}_full -f python bench.py --no-e2e --no-cpu --steps 2 --warmup 3 \
    > $out/${tag}_ncu_full.log 2>&1
# 5. the dual regression's NNLS kernel (stage 1 and stage 2 launches) on a C3-shaped subject
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nnls_smem_kernel -c 2 \
    -o $out/${tag}_nnls_full -f python tests/gpu_dr_probe.py C3 1 > $out/${tag}_ncu_nnls.log 2>&1
ls -la $out | tail -8
