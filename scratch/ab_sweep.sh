#!/bin/bash
# A/B: new lib vs scratch/lib_old.so, sweeps at N=8-shard sizes and at C3
set -x
timeout 300 python -m pytest tests/test_nmf_gpu.py tests/test_edge_gpu.py -m gpu -x -q 2>&1 | tail -3
run() {
  NFB_PHASE_TIMING=1 timeout 200 python bench.py --shape 7426,13750,200 --steps 20 --warmup 3 --no-e2e --no-cpu --no-dr 2>&1 | grep -E "phases over 20" | tail -1 | cut -c1-700
  NFB_PHASE_TIMING=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-dr 2>&1 | grep -E "phases over 10|\"value\"" | cut -c1-700
}
echo NEW; run
cp nfact_b200/libnfact_b200.so scratch/lib_new.so; cp scratch/lib_old.so nfact_b200/libnfact_b200.so
echo OLD; run
cp scratch/lib_new.so nfact_b200/libnfact_b200.so
