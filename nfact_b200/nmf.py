"""NMF solve -- host-side mirror of the reference's decomposition seam, arithmetic on the GPU.

Same names, arguments and error behaviour as
  NFACT/decomp/decomposition/sso/nmfsso.py:32-59          nmf_decomp(parameters, fdt_matrix, W=None, H=None)
  NFACT/decomp/decomposition/matrix_decomposition.py:74-107   get_parameters(parameters, algo, n_components)
  NFACT/config/nfact_config_functions.py:337-356          the hyper-parameter schema (= sklearn NMF's signature)
The solver itself (sklearn/decomposition/_nmf.py: _fit_coordinate_descent :399-518,
_fit_multiplicative_update :726-897, _initialize_nmf :214-366) runs in libnfact_b200.
"""
from __future__ import annotations

import numpy as np

from . import device
from .device import DeviceMatrix, NmfState, make_params
from .utils import error_and_exit

# sklearn 1.9 NMF.__init__ signature minus n_components (what ``nfact_config -D`` dumps)
NMF_DEFAULTS = {
    "init": None, "solver": "cd", "beta_loss": "frobenius", "tol": 1e-4, "max_iter": 200,
    "random_state": None, "alpha_W": 0.0, "alpha_H": "same", "l1_ratio": 0.0, "verbose": 0, "shuffle": False,
}


def create_combined_algo_dict() -> dict:
    """NFACT/config/nfact_config_functions.py:337-356 (NMF block)."""
    return {"nmf": dict(NMF_DEFAULTS)}


def get_parameters(parameters: dict, algo: str, n_components: int) -> dict:
    """NFACT/decomp/decomposition/matrix_decomposition.py:74-107."""
    if parameters:
        parameters["n_components"] = n_components
        return parameters
    if algo != "nmf":
        raise ValueError("nfact_b200 implements the NMF decomposition path only")
    parameters = create_combined_algo_dict()[algo]
    parameters["n_components"] = n_components
    parameters["alpha_W"] = 0.1
    parameters["init"] = "nndsvd"
    parameters["random_state"] = None
    parameters["l1_ratio"] = 1
    return parameters


def compute_regularization(n_samples: int, n_features: int, alpha_W, alpha_H, l1_ratio):
    """sklearn/decomposition/_nmf.py:1249-1259 (float64 scalars)."""
    alpha_H = alpha_W if alpha_H == "same" else alpha_H
    return (n_features * alpha_W * l1_ratio, n_samples * alpha_H * l1_ratio,
            n_features * alpha_W * (1.0 - l1_ratio), n_samples * alpha_H * (1.0 - l1_ratio))


def handle_comm(handle):
    """Collectives of the handle's row-sharded world (``nndsvd.Comm`` over torch.distributed, which is also what
    brought up the library's NCCL communicator); the identity when the handle is not sharded."""
    from .nndsvd import Comm
    if getattr(handle, "world", 1) <= 1:
        return Comm()
    return Comm(handle.rank, handle.world, getattr(handle, "group", None))


def global_rows(handle, m_local: int) -> int:
    """Seed rows of the whole matrix when X is row-sharded over the handle's NCCL world (one process per
    GPU, torch.distributed as the rendezvous); the local count otherwise."""
    if getattr(handle, "world", 1) <= 1:
        return int(m_local)
    import torch
    return handle_comm(handle).row_layout(m_local, torch.device("cuda", handle.device))[1]


def _check_params(p: dict) -> dict:
    unknown = set(p) - set(NMF_DEFAULTS) - {"n_components"}
    if unknown:
        raise TypeError(f"NMF.__init__() got an unexpected keyword argument '{sorted(unknown)[0]}'")
    full = dict(NMF_DEFAULTS)
    full.update(p)
    if full["beta_loss"] not in ("frobenius", 2, 2.0):
        raise ValueError("nfact_b200 implements beta_loss='frobenius' only (no CPU fallback for other losses)")
    if full["solver"] not in ("cd", "mu"):
        raise ValueError(f"The 'solver' parameter of NMF must be a str among {{'cd', 'mu'}}. Got {full['solver']!r}")
    if full["shuffle"]:
        raise ValueError("shuffle=True is not implemented on the device (NFACT never sets it)")
    if not isinstance(full["n_components"], (int, np.integer)) or full["n_components"] < 1:
        raise ValueError("n_components must be a positive integer")
    return full


def _random_state(seed):
    if seed is None or isinstance(seed, (int, np.integer)):
        return np.random.RandomState(seed)
    if isinstance(seed, np.random.RandomState):
        return seed
    raise ValueError(f"{seed!r} cannot be used to seed a numpy.random.RandomState instance")


def _shared_random_state(seed, comm):
    """RandomState every rank agrees on: a row-sharded initialisation draws ONE stream and slices it by global
    seed row, so ``random_state=None`` (NFACT's default) needs a seed picked by rank 0."""
    if comm.world > 1 and not isinstance(seed, (int, np.integer)):
        import torch.distributed as dist
        if isinstance(seed, np.random.RandomState):
            box = [int(seed.randint(0, 2 ** 31 - 1))]
        elif seed is None:
            box = [int(np.random.RandomState().randint(0, 2 ** 31 - 1))]
        else:
            box = [seed]
        dist.broadcast_object_list(box, src=0, group=comm.group)
        seed = box[0]
    return _random_state(seed)


def initialize_nmf(xdev: DeviceMatrix, x_mean: float, k: int, init, random_state, comm=None):
    """sklearn/decomposition/_nmf.py:214-366 for init in {random, nndsvd, nndsvda, nndsvdar}.

    The random draws come from numpy's RandomState exactly as in sklearn (so a fixed ``random_state``
    gives sklearn's starting point); every product with X runs on the device.

    ``comm`` (``nndsvd.Comm``, world > 1): ``xdev`` holds this rank's seed rows and ``x_mean`` is the mean of
    the WHOLE matrix; the returned W has this rank's rows of the unsharded initialisation, H is the same on
    every rank.  Every rank draws the full random stream and slices it by global seed row.
    """
    from .nndsvd import Comm, nndsvd_device
    comm = comm or Comm()
    m_local, n = xdev.shape
    row0, m = 0, m_local
    if comm.world > 1:
        import torch
        row0, m = comm.row_layout(m_local, torch.device("cuda", xdev.handle.device))
    if init is None:
        init = "nndsvda" if k <= min(m, n) else "random"
    if init != "random" and k > min(m, n):
        raise ValueError(f"init = '{init}' can only be used when n_components <= min(n_samples, n_features)")
    rng = _shared_random_state(random_state, comm)
    if init == "random":
        avg = np.sqrt(x_mean / k)
        H = avg * rng.standard_normal(size=(k, n)).astype(np.float32, copy=False)
        W = avg * rng.standard_normal(size=(m, k)).astype(np.float32, copy=False)
        W = np.ascontiguousarray(W[row0:row0 + m_local])
        return np.abs(W, out=W), np.abs(H, out=H)
    if init not in ("nndsvd", "nndsvda", "nndsvdar"):
        raise ValueError(f"Invalid init parameter: got {init!r} instead of one of "
                         "(None, 'random', 'nndsvd', 'nndsvda', 'nndsvdar')")
    W, H = nndsvd_device(xdev, k, rng, comm=comm)
    if init == "nndsvda":
        W[W == 0] = x_mean
        H[H == 0] = x_mean
    elif init == "nndsvdar":
        # sklearn fills the zeros of W (row-major), then of H, from one stream: a rank's zeros of W are a
        # contiguous run of that sequence
        nz_w = int((W == 0).sum())
        before, total = 0, nz_w
        if comm.world > 1:
            import torch
            counts = comm.allgather(torch.tensor([nz_w], dtype=torch.int64,
                                                 device=torch.device("cuda", xdev.handle.device)))
            counts = [int(c.item()) for c in counts]
            before, total = sum(counts[:comm.rank]), sum(counts)
        draws = rng.standard_normal(size=total)
        W[W == 0] = abs(x_mean * draws[before:before + nz_w] / 100)
        H[H == 0] = abs(x_mean * rng.standard_normal(size=int((H == 0).sum())) / 100)
    return W, H


def _check_custom(W, H, m: int, n: int, k: int) -> None:
    """sklearn/decomposition/_nmf.py:1186-1203 (_check_w_h -> _check_init) for ``init='custom'``."""
    for name, arr, shape in (("H", H, (k, n)), ("W", W, (m, k))):
        arr = np.asarray(arr)
        if arr.shape != shape:
            raise ValueError(f"Array with wrong shape passed to NMF (input {name}). "
                             f"Expected {shape}, but got {arr.shape} ")
        if arr.size and arr.min() < 0:
            raise ValueError(f"Negative values in data passed to NMF (input {name})")
        if arr.size and arr.max() == 0:
            raise ValueError(f"Array passed to NMF (input {name}) is full of zeros.")


class NMF:
    """The slice of sklearn.decomposition.NMF that NFACT touches (nmfsso.py:54-59):
    ``NMF(**parameters).fit_transform(X, W=W, H=H)``, ``components_``, ``n_iter_``,
    ``reconstruction_err_``."""

    def __init__(self, n_components=None, **kwargs):
        self.params = _check_params(dict(kwargs, n_components=n_components))

    def fit_transform(self, X, y=None, W=None, H=None):
        import os
        import time
        p = self.params
        k = int(p["n_components"])
        marks = [("start", time.perf_counter())] if os.environ.get("NFB_HOST_TIMING") else None

        def mark(name):
            if marks is not None:
                marks.append((name, time.perf_counter()))
        if isinstance(X, DeviceMatrix):
            xdev, x_mean, owns = X, None, False
        else:
            X = np.asarray(X)
            if X.ndim != 2:
                raise ValueError(f"Expected 2D array, got {X.ndim}D array instead")
            if X.dtype not in (np.float32, np.float64):
                X = X.astype(np.float64)
            from . import multigpu
            if multigpu.active():
                # torchrun drop-in, rank 0: a host matrix is solved row-sharded over all GPUs (multigpu.py)
                if p["init"] == "custom":
                    _check_custom(W, H, X.shape[0], X.shape[1], k)
                Wf, Hf, self.n_iter_, self.reconstruction_err_ = multigpu.nmf_fit_sharded(p, X, W, H)
                self.components_, self.n_components_ = Hf, Hf.shape[0]
                return Wf
            # the mean is only needed by the non-custom initialisations; for a large matrix a host pass costs
            # seconds, so it is taken from the float64 sum the device accumulates while splitting X
            x_mean = float(X.mean()) if p["init"] != "custom" and X.nbytes <= (256 << 20) else None
            xdev, owns = DeviceMatrix.from_host(X), True
        mark("upload+split X")
        try:
            if xdev.has_negative:     # sklearn check_non_negative, evaluated on the device during ingest
                raise ValueError("Negative values in data passed to NMF (input X)")
            m, n = xdev.shape
            if p["init"] == "custom":
                _check_custom(W, H, m, n, k)
                W0 = np.array(W, dtype=np.float32, order="C")
                H0 = np.array(H, dtype=np.float32, order="C")
            else:
                comm = handle_comm(xdev.handle)
                if comm.world > 1:      # seed-row shard: mean of the whole matrix from the per-rank float64 sums
                    import torch
                    tot = torch.tensor([xdev.mean * m * n, float(m)], dtype=torch.float64,
                                       device=torch.device("cuda", xdev.handle.device))
                    tot = comm.allreduce(tot)
                    x_mean = float(tot[0]) / (float(tot[1]) * n)
                elif x_mean is None:    # resident matrix: float64 mean accumulated during the split
                    x_mean = xdev.mean
                W0, H0 = initialize_nmf(xdev, x_mean, k, p["init"], p["random_state"], comm)
            # row-sharded runs: sklearn's n_samples is the number of seeds of the whole matrix
            regs = compute_regularization(global_rows(xdev.handle, m), n, p["alpha_W"], p["alpha_H"], p["l1_ratio"])
            mark("init / checks")
            state = NmfState(xdev, k, make_params(p["solver"], p["max_iter"], p["tol"], regs, p["verbose"]))
            mark("state")
            try:
                state.set_factors(W0, H0)
                mark("factor upload")
                self.n_iter_, self.reconstruction_err_ = state.run()
                mark("solve")
                Wf, Hf = state.get_factors()
                mark("download")
            finally:
                state.free()
        finally:
            if owns:
                xdev.free()
        mark("free")
        if marks is not None:
            import sys
            print("nfact_b200 host timing: " + "  ".join(f"{b[0]}={b[1] - a[1]:.3f}s" for a, b in zip(marks, marks[1:])),
                  file=sys.stderr)
        self.components_ = Hf
        self.n_components_ = Hf.shape[0]
        return Wf


def nmf_decomp(parameters: dict, fdt_matrix, W=None, H=None) -> dict:
    """NFACT/decomp/decomposition/sso/nmfsso.py:32-59.

    ``fdt_matrix`` is the float32 [m, n] connectivity matrix (a numpy array like the reference, or a
    ``DeviceMatrix`` that is already resident); ``W``/``H`` given => ``init='custom'``.
    Returns {"grey_components": W [m,k], "white_components": H [k,n]} as float32 numpy arrays.
    """
    if W is not None and H is not None:
        parameters = parameters.copy()
        parameters["init"] = "custom"
    grey_matter = components = None
    try:
        decomp = NMF(**parameters)
        grey_matter = decomp.fit_transform(fdt_matrix, W=W, H=H)
        components = decomp.components_
    except Exception as e:
        error_and_exit(False, f"Unable to perform NMF due to {e}")
    return {"grey_components": grey_matter, "white_components": components}
