"""Probe: the nfact_dr subject loop (stream_subjects) on C3-shaped subjects given as page-locked COO triplets:
serial (ingest then regress) against prefetching (subject i + 1 is ingested while subject i is regressed)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402
import nfact_b200  # noqa: E402
from bench import WORKLOADS, synth_rows_device  # noqa: E402
from nfact_b200.device import ingest_coo_resident  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
n_sub = int(sys.argv[2]) if len(sys.argv) > 2 else 6
m, n, k, density = WORKLOADS[name]
dev = torch.device("cuda", 0)
x = synth_rows_device(torch, dev, 0, m, n, k, density)
idx = x.nonzero()
rows = idx[:, 0].to(torch.int32).cpu().pin_memory().numpy()
cols = idx[:, 1].to(torch.int32).cpu().pin_memory().numpy()
vals = x[idx[:, 0], idx[:, 1]].to(torch.float64).cpu().pin_memory().numpy()
del x, idx
torch.cuda.empty_cache()
rng = np.random.default_rng(0)
comps = {"grey_components": rng.gamma(0.3, 1.0, (m, k)) * (rng.random((m, k)) < 0.3)}


def load(index, handle):
    return ingest_coo_resident([(rows, cols, vals, 1.0)], m, n, False, handle)


for prefetch in (False, True, True, True):
    t0 = time.perf_counter()
    marks = []
    for i, res in nfact_b200.stream_subjects(load, n_sub, comps, prefetch=prefetch):
        marks.append(time.perf_counter() - t0)
        del res
    steady = (marks[-1] - marks[0]) / (n_sub - 1)
    print(f"prefetch={prefetch}: {n_sub} subjects in {marks[-1]:.3f} s, steady state {steady * 1e3:.1f} ms/subject "
          f"({1 / steady:.2f} subjects/s)", flush=True)
