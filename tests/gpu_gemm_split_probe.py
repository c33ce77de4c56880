"""Probe: W'X-type GEMM (A read MN-major) at C3 size, N = 208, as a function of the split-K factor."""
import ctypes
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402
from bench import WORKLOADS, synth_rows_device  # noqa: E402
from nfact_b200 import _lib  # noqa: E402
from nfact_b200.device import get_handle  # noqa: E402

m, n, k, density = WORKLOADS["C3"]
dev = torch.device("cuda", 0)
x = synth_rows_device(torch, dev, 0, m, n, k, density)
h = get_handle()
w = torch.rand((k, m), device=dev)
hh = torch.rand((k, n), device=dev)
o = torch.empty((k, max(m, n)), device=dev)
torch.cuda.synchronize()
for splits in [int(a) for a in sys.argv[1:]] or [1, 2, 3, 6]:
    for trans in (1, 0):
        b = w if trans else hh
        ld = n if trans else m
        _lib.check(h.lib.nfb_test_gemm(h.ptr, ctypes.c_void_p(x.data_ptr()), m, n, trans, ctypes.c_void_p(b.data_ptr()), k,
                                       ctypes.c_void_p(o.data_ptr()), ld, 2, splits))
