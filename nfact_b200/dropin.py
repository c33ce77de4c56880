"""Run the UNMODIFIED NFACT command-line tools on the B200 path.

NFACT has no plugin layer: its hot path leaves Python at a handful of module-level functions (SURVEY.md
section 8b).  ``install()`` rebinds exactly those names -- in the module that defines each one and in every
NFACT module that ``from``-imported it -- to the ``nfact_b200`` functions of the same name and signature, so
``nfact_decomp`` / ``nfact_dr`` keep their flags, ``nfact_config`` hyper-parameters, logging, image writers and
output files and only the arithmetic moves to the GPU::

    python -m nfact_b200.dropin nfact_decomp -l sub_list -s seeds... -o out -d 200
    python -m nfact_b200.dropin nfact_dr -l sub_list -n out/nfact_decomp -a NMF -o out

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

_ENTRY_POINTS = {   # pyproject.toml:41-48
    "nfact_decomp": ("NFACT.decomp.__main__", "nfact_decomp_main"),
    "nfact_dr": ("NFACT.dual_reg.__main__", "nfact_dr_main"),
}

_originals: dict = {}


def _replacement(name: str):
    import nfact_b200
    from nfact_b200 import nmfsso
    if hasattr(nfact_b200, name):
        return getattr(nfact_b200, name)
    return getattr(nmfsso, name)


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
    sys.argv = [tool] + argv[1:]                   # the reference parses sys.argv itself
    getattr(module, func)()


if __name__ == "__main__":
    main()
