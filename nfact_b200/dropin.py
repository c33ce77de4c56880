"""Run the UNMODIFIED NFACT command-line tools on the B200 path.

NFACT has no plugin layer: its hot path leaves Python at a handful of module-level functions (SURVEY.md
section 8b).  ``install()`` rebinds exactly those names -- in the module that defines each one and in every
NFACT module that ``from``-imported it -- to the ``nfact_b200`` functions of the same name and signature, so
``nfact_decomp`` / ``nfact_dr`` keep their flags, ``nfact_config`` hyper-parameters, logging, image writers and
output files and only the arithmetic moves to the GPU::

    python -m nfact_b200.dropin nfact_decomp -l sub_list -s seeds... -o out -d 200
    python -m nfact_b200.dropin nfact_dr -l sub_list -n out/nfact_decomp -a NMF -o out

or on all GPUs of a node (rank 0 runs the tool, the other ranks serve its solves: ``nfact_b200/multigpu.py``)::

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 -m nfact_b200.dropin nfact_decomp ...

or, from Python, ``import nfact_b200.dropin; nfact_b200.dropin.install()`` before NFACT's mains are imported
or called.  There is no CPU fallback behind the patched names: without ``libnfact_b200.so`` or a Blackwell GPU
they raise, which NFACT reports through its own ``error_and_exit``.  ICA (``--algo ICA``) is out of scope and
keeps the reference's functions untouched.
"""
from __future__ import annotations

import importlib
import sys

# module that DEFINES the name  ->  names rebound there (value: attribute of nfact_b200 / nfact_b200.nmfsso)
_DEFINITIONS = {
    # ingest a1-a5 (NFACT/base/matrix_handling.py:93-181, decomp/decomposition/matrix_handling.py:12-100)
    "NFACT.base.matrix_handling": (
        "read_in_matrix", "load_fdt_matrix", "save_matrix",     # (load_single_matrix keeps its scipy return type)
        # post-processing a14 (:45-68, :184-277)
        "thresholding", "threshold_grey_components", "thresholding_components", "normalise_components"),
    "NFACT.decomp.decomposition.matrix_handling": ("process_fdt_matrix2", "avg_fdt", "load_previous_matrix"),
    # solve a7, a13 (nmfsso.py:32-59, 62-238, 322-360; sso_functions.py:29-68)
    "NFACT.decomp.decomposition.sso.nmfsso": ("nmf_decomp", "nmf_sso", "NMFsso"),
    "NFACT.decomp.decomposition.sso.sso_functions": ("compute_similairty_matrix",),
    # hyper-parameters a6 (matrix_decomposition.py:74-107)
    "NFACT.decomp.decomposition.matrix_decomposition": ("get_parameters",),
    # winner-takes-all maps f3 (pipes/image_handling.py:216-245; winner_takes_all :186-187 looks the name up here)
    "NFACT.decomp.pipes.image_handling": ("create_wta_map",),
    # dual regression a15 (dual_regression_methods.py:8-42, 92-116)
    "NFACT.dual_reg.dual_regression.dual_regression_methods": ("nmf_dual_regression", "run_decomp"),
}

# modules that took copies of those names with ``from ... import`` (reference import lines cited)
_IMPORTERS = {
    "NFACT.decomp.decomposition.matrix_handling": ("load_fdt_matrix",),                                 # :6
    "NFACT.decomp.decomposition.sso.nmfsso": ("thresholding", "compute_similairty_matrix"),              # :1-15
    "NFACT.decomp.decomposition.matrix_decomposition": ("normalise_components", "nmf_sso"),              # :5-7
    "NFACT.decomp.__main__": ("save_matrix", "process_fdt_matrix2", "load_previous_matrix",              # :5, :24-37
                              "get_parameters", "thresholding_components"),
    "NFACT.dual_reg.dual_regression.run_dual_regression": ("load_fdt_matrix", "nmf_dual_regression",     # :2-10
                                                           "run_decomp", "thresholding_components",
                                                           "normalise_components"),
}

_DR_LOOP = ("NFACT.dual_reg.dual_regression.set_up_dual_regression", "run_locally")
_DR_LOOP_IMPORTERS = ("NFACT.dual_reg.__main__",)                                                        # :2-5

_ENTRY_POINTS = {   # pyproject.toml:41-48
    "nfact_decomp": ("NFACT.decomp.__main__", "nfact_decomp_main"),
    "nfact_dr": ("NFACT.dual_reg.__main__", "nfact_dr_main"),
}

_originals: dict = {}


def get_parameters(parameters: dict, algo: str, n_components: int) -> dict:
    """Drop-in for NFACT/decomp/decomposition/matrix_decomposition.py:74-107.  The NMF block is nfact_b200's mirror
    (same defaults); every other algorithm (``--algo ICA``) goes to the reference's own function untouched --
    ICA is out of this package's scope and must keep working under the drop-in."""
    if str(algo).lower() == "nmf" or parameters:
        import nfact_b200
        return nfact_b200.get_parameters(parameters, algo, n_components)
    original = _originals.get(("NFACT.decomp.decomposition.matrix_decomposition", "get_parameters"))
    if original is None:
        raise ValueError("nfact_b200 implements the NMF decomposition path only")
    return original(parameters, algo, n_components)


def _replacement(name: str):
    import nfact_b200
    from nfact_b200 import nmfsso
    if name == "get_parameters":
        return get_parameters
    if hasattr(nfact_b200, name):
        return getattr(nfact_b200, name)
    return getattr(nmfsso, name)


def run_locally(args: dict, paths: dict) -> None:
    """Drop-in for ``run_locally(args, paths)``, NFACT/dual_reg/dual_regression/set_up_dual_regression.py:186-235:
    the subject loop of ``nfact_dr``.  The reference loads, regresses, post-processes and saves one subject after the
    other through ``dual_regression_pipeline`` (run_dual_regression.py:66-199); here the same steps run through
    ``nfact_b200.run_locally``: the next subject's ``fdt_matrix2.dot[.lz4]`` is decoded, parsed and ingested straight
    into HBM while the current one is regressed, the group components are uploaded once, thresholding / z-scoring
    happen on the device, and every result goes to the reference's own ``save_dual_regression_images`` with the
    reference's arguments.  ICA keeps the reference's loop."""
    import os
    original = _originals.get(_DR_LOOP)
    if str(args["algo"]).lower() != "nmf":
        if original is None:
            raise RuntimeError("nfact_b200 only replaces the NMF dual regression")
        return original(args, paths)
    import nfact_b200
    from nfact_b200.utils import colours, error_and_exit, nprint
    drf = importlib.import_module("NFACT.dual_reg.nfact_dr_functions")
    col = colours()
    nprint(f"{col['pink']}Running:{col['reset']} Locally")
    nprint(f"{col['pink']}DR Method:{col['reset']} Non-negative Regression")
    nprint(f"{col['pink']}Obtaining:{col['reset']} Components")
    try:
        components = drf.get_group_level_components(paths["component_path"], paths["group_average_path"],
                                                    args["seeds"], args["roi"])
    except Exception as e:
        error_and_exit(False, f"Unable to find components due to {e}")
    spec = {"out_dir": os.path.join(args["outdir"], "nfact_dr"), "seeds": args["seeds"],
            "algo": str(args["algo"]).upper(), "roi": args["roi"], "cifti": args["cifti"]}
    read_ahead = int(os.environ.get("NFB_READ_AHEAD", "0"))
    from nfact_b200 import multigpu
    if multigpu.active():
        # torchrun: subjects rank, rank + world, ... per GPU; every rank writes its own subjects' images
        nprint(f"{col['pink']}GPUs:{col['reset']} {multigpu.context().world}")
        multigpu.dr_distributed(list(args["ptxdir"]), components, args["seeds"], int(args["threshold"]),
                                bool(args["normalise"]), spec, read_ahead)
        return None
    # NFB_READ_AHEAD=N: N host threads read / decompress the next subjects' files ahead of the GPU (run_locally)
    nfact_b200.run_locally(list(args["ptxdir"]), components, args["seeds"], threshold=int(args["threshold"]),
                           normalise=bool(args["normalise"]), save=make_dr_saver(spec), read_ahead=read_ahead)
    return None


def make_dr_saver(spec: dict):
    """``save(subject_dir, dr_results)`` that hands a subject's result to the reference's own
    ``save_dual_regression_images`` with the reference's arguments (run_dual_regression.py:181-196)."""
    from nfact_b200.utils import colours, error_and_exit, nprint
    drf = importlib.import_module("NFACT.dual_reg.nfact_dr_functions")
    col = colours()

    def save(subject_dir, dr_results):
        sub_id = drf.get_subject_id(subject_dir)
        nprint(f"{col['pink']}Saving{col['reset']}: Components ({sub_id})", to_flush=True)
        try:
            drf.save_dual_regression_images(dr_results, spec["out_dir"], spec["seeds"], spec["algo"],
                                            dr_results["white_components"].shape[0], sub_id, subject_dir,
                                            spec["roi"], spec["cifti"])
        except Exception as e:
            error_and_exit(False, f"Unable to save images due to {e}")
        nprint(f"{col['pink']}Completed{col['reset']}: {sub_id}", to_flush=True)
    return save


def install(strict: bool = False) -> list:
    """Rebind the hot-path functions of an importable NFACT to their nfact_b200 counterparts.

    Returns the list of ``"module.name"`` strings that were rebound.  A module that cannot be imported (an
    optional dependency of the reference such as nibabel is missing) is skipped unless ``strict``; a name a
    module does not have is skipped the same way.  Idempotent."""
    from ._lib import load
    load()                                   # fail here, loudly, when the CUDA library is missing
    patched = []
    for table in (_DEFINITIONS, _IMPORTERS):
        for modname, names in table.items():
            try:
                module = importlib.import_module(modname)
            except Exception:
                if strict:
                    raise
                continue
            for name in names:
                if not hasattr(module, name):
                    if strict:
                        raise AttributeError(f"{modname} has no attribute {name!r}")
                    continue
                new = _replacement(name)
                old = getattr(module, name)
                if old is new:
                    continue
                _originals.setdefault((modname, name), old)
                setattr(module, name, new)
                patched.append(f"{modname}.{name}")
    # the subject loop of nfact_dr (not a same-named nfact_b200 function: it takes the reference's args / paths dicts)
    for modname in (_DR_LOOP[0],) + _DR_LOOP_IMPORTERS:
        try:
            module = importlib.import_module(modname)
        except Exception:
            if strict:
                raise
            continue
        old = getattr(module, _DR_LOOP[1], None)
        if old is None or old is run_locally:
            continue
        _originals.setdefault((modname, _DR_LOOP[1]), old)
        setattr(module, _DR_LOOP[1], run_locally)
        patched.append(f"{modname}.{_DR_LOOP[1]}")
    return patched


def uninstall() -> None:
    """Put the reference's own functions back (tests)."""
    for (modname, name), old in list(_originals.items()):
        module = sys.modules.get(modname)
        if module is not None:
            setattr(module, name, old)
    _originals.clear()


def installed() -> list:
    return sorted(f"{m}.{n}" for m, n in _originals)


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in _ENTRY_POINTS:
        raise SystemExit("usage: python -m nfact_b200.dropin {nfact_decomp|nfact_dr} [the tool's own arguments]")
    tool = argv[0]
    install()
    modname, func = _ENTRY_POINTS[tool]
    module = importlib.import_module(modname)      # from-imports inside now bind the rebound functions
    install()                                      # and anything the main module imported under its own name
    # torchrun --nproc-per-node N -m nfact_b200.dropin ...: rank 0 runs the tool, the other ranks serve its
    # solves (nfact_b200/multigpu.py); a plain `python -m` launch is the single-GPU tool
    from nfact_b200 import multigpu
    ctx = multigpu.init_from_env()
    if ctx is not None and ctx.rank != 0:
        multigpu.serve()
        return
    sys.argv = [tool] + argv[1:]                   # the reference parses sys.argv itself
    try:
        getattr(module, func)()
    finally:
        if ctx is not None:
            multigpu.shutdown()


if __name__ == "__main__":
    main()
