"""Probe: systematic (signed) error of the TMEM fp32 accumulation chain as a function of chain length."""
import ctypes
import sys

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402
import nfact_b200  # noqa: E402
from nfact_b200._lib import check  # noqa: E402

h = nfact_b200.get_handle()
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev)
gen.manual_seed(0)
rows, K, nb = 512, 110000, 200
a = torch._standard_gamma(torch.full((rows, K), 0.3, device=dev), generator=gen) * 300
a = (a * (torch.rand((rows, K), device=dev, generator=gen) < 0.05)).float().contiguous()
b = torch._standard_gamma(torch.full((nb, K), 0.5, device=dev), generator=gen).float().contiguous()
ref = (b.double() @ a.double().t())
for splits in (1, 6, 24, 96, 384):
    out = torch.zeros((nb, rows), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    check(h.lib.nfb_test_gemm(h.ptr, ctypes.c_void_p(a.data_ptr()), rows, K, 0, ctypes.c_void_p(b.data_ptr()), nb,
                              ctypes.c_void_p(out.data_ptr()), rows, 2, splits))
    rel = ((out.double() - ref) / ref)
    print(f"K={K} splits={splits:4d} k-blocks/chain={K/64/splits:7.1f}  mean rel={rel.mean().item():+.3e} "
          f"max|rel|={rel.abs().max().item():.3e}", flush=True)
s32 = (b @ a.t())
rel = ((s32.double() - ref) / ref)
print(f"torch fp32 matmul (TF32 off={not torch.backends.cuda.matmul.allow_tf32}) mean rel={rel.mean().item():+.3e} "
      f"max|rel|={rel.abs().max().item():.3e}")
