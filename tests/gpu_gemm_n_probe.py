"""Probe: device time of the X-streaming GEMM as a function of the UMMA N (number of components) at C3 size."""
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402
from bench import WORKLOADS, synth_rows_device  # noqa: E402
from nfact_b200.device import DeviceMatrix  # noqa: E402

m, n, k, density = WORKLOADS["C3"]
dev = torch.device("cuda", 0)
x = synth_rows_device(torch, dev, 0, m, n, k, density)
torch.cuda.synchronize()
xmat = DeviceMatrix.from_device_ptr(x.data_ptr(), m, n, n)
del x
torch.cuda.empty_cache()
for nb in [int(a) for a in sys.argv[1:]] or [128, 160, 192, 208, 224, 256]:
    h = torch.rand((nb, n), device=dev)
    w = torch.rand((nb, m), device=dev)
    o1 = torch.empty((nb, m), device=dev)
    o2 = torch.empty((nb, n), device=dev)
    torch.cuda.synchronize()
    for rep in range(3):
        xmat.matmul_dev(False, h.data_ptr(), nb, o1.data_ptr(), m)
        xmat.matmul_dev(True, w.data_ptr(), nb, o2.data_ptr(), n)
