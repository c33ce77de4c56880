"""Turn the scratch outputs of profiles/collect.sh (gpurun_out/<tag>_*) into the tracked summaries:

    python profiles/summarise.py r01b

  profiles/<tag>_launches.csv           the ncu launch list (gpu__time_duration.sum, --clock-control none) of
                                        `bench.py --no-e2e --no-cpu --steps 2 --warmup 3`, own kernels only
  profiles/<tag>_launches_summary.csv   per-kernel launches / ms per iteration / share of the iteration (last 2
                                        of the 5 iterations; under ncu every launch is cold-cache and serialised)
  profiles/<tag>_ncu_full.csv           selected `--set full` metrics of the captured GEMM / sweep launches
  profiles/<tag>_bench.json             the bench line of the same run
"""
import collections
import csv
import os
import re
import shutil
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
src, dst = "gpurun_out", "profiles"


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"(nfb::)?<?unnamed>::", "", name)
    return re.sub(r"\(.*$", "", name)


rows = list(csv.reader(l for l in open(os.path.join(src, f"{tag}_launches.csv")) if l.startswith('"')))
hdr, data = rows[0], rows[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
launches = [(short(r[ki]), float(r[vi]) / 1e6) for r in data]          # ms
with open(os.path.join(dst, f"{tag}_launches.csv"), "w") as fh:
    fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:nfb:: of: "
             "python bench.py --no-e2e --no-cpu --steps 2 --warmup 3 (C3, CD, 1 B200)\nindex,kernel,ms\n")
    for i, (k, ms) in enumerate(launches):
        fh.write(f'{i},"{k}",{ms:.6f}\n')
# iterations start at every X H' GEMM (gemm_split_kernel<0, ...> with the full-size grid): take the last two
starts = [i for i, (k, ms) in enumerate(launches) if k.startswith("gemm_split_kernel<0") and ms > 1.0]
# two complete iterations: from the third-last iteration start up to the last one
per_iter = starts[-1] - starts[-2]
it = launches[starts[-3]:starts[-1]]
agg = collections.OrderedDict()
for k, ms in it:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ms
total = sum(a[1] for a in agg.values())
with open(os.path.join(dst, f"{tag}_launches_summary.csv"), "w") as fh:
    fh.write(f"# two full CD iterations at C3 (59412x110000, k=200), 1 B200, under ncu (cold-cache, serialised): "
             f"{total / 2:.3f} ms per iteration in {per_iter} launches\nkernel,launches_per_iter,ms_per_iter,share\n")
    for k, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write(f'"{k}",{cnt / 2:g},{ms / 2:.4f},{ms / total:.4f}\n')
print(open(os.path.join(dst, f"{tag}_launches_summary.csv")).read())

rep = os.path.join(src, f"{tag}_full.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h = rr[0]
    want = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum"]
    idx = [h.index(w) for w in want if w in h]
    with open(os.path.join(dst, f"{tag}_ncu_full.csv"), "w") as fh:
        w = csv.writer(fh)
        fh.write("# ncu --set full --clock-control none --import-source on -k regex:'gemm_split_kernel|cd_sweep_tile' "
                 "--launch-skip 12 -c 6 of the same bench command; selected metrics (row 2 = units)\n")
        for r in rr:
            w.writerow([short(r[i]) if h[i] == "Kernel Name" and r[i] else r[i] for i in idx])
    print(open(os.path.join(dst, f"{tag}_ncu_full.csv")).read()[:3000])
if os.path.exists(os.path.join(src, f"{tag}_bench.json")):
    shutil.copy(os.path.join(src, f"{tag}_bench.json"), os.path.join(dst, f"{tag}_bench_C3_cd.json"))
nnls = os.path.join(src, f"{tag}_nnls_full.ncu-rep")
if os.path.exists(nnls):
    subprocess.run([sys.executable, os.path.join(dst, "extract_full.py"), nnls, os.path.join(dst, f"{tag}_ncu_full_nnls_C3.csv"),
                    "ncu --set full --clock-control none --import-source on -k regex:nnls_smem_kernel -c 2 python "
                    "tests/gpu_dr_probe.py C3 1 (stage 1: 110000 columns, stage 2: 59412 rows, k=200)"],
                   stdout=subprocess.DEVNULL)
