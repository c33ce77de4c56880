"""nfact_b200.dropin: the reference's seams are rebound to functions of the same name and calling convention.

The binding test imports the reference from /root/reference when it exists (build container); on the GPU box it
does not exist and only the table checks run.  No compute: nothing here needs a GPU.
"""
import importlib
import inspect
import os
import sys
import types

import numpy as np
import pytest

import nfact_b200
from nfact_b200 import dropin

REF = "/root/reference"


class _Stub(types.ModuleType):
    """Stands in for an optional dependency of the reference that is not installed here (nibabel, lz4, ...)."""
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        child = _Stub(f"{self.__name__}.{name}")
        setattr(self, name, child)
        return child

    def __call__(self, *args, **kwargs):
        return self


@pytest.fixture
def reference():
    if not os.path.isdir(os.path.join(REF, "NFACT")):
        pytest.skip("reference checkout not present (GPU box)")
    added = []
    for name in ("lz4", "lz4.frame", "nibabel", "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "seaborn",
                 "file_tree", "networkx", "fsl", "fsl.wrappers", "fsl_sub"):
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _Stub(name)
            added.append(name)
    sys.path.insert(0, REF)
    yield
    dropin.uninstall()
    sys.path.remove(REF)
    for name in added:
        sys.modules.pop(name, None)
    for name in [n for n in sys.modules if n == "NFACT" or n.startswith("NFACT.")]:
        del sys.modules[name]


def test_every_rebound_name_exists_in_the_product():
    for table in (dropin._DEFINITIONS, dropin._IMPORTERS):
        for names in table.values():
            for name in names:
                assert callable(dropin._replacement(name)), name
    # the seams of SURVEY.md section 8b are all in the table
    rebound = {n for names in dropin._DEFINITIONS.values() for n in names}
    assert {"process_fdt_matrix2", "load_fdt_matrix", "nmf_decomp", "nmf_dual_regression", "run_decomp",
            "get_parameters"} <= rebound


def test_install_rebinds_the_reference_seams(reference):
    ref_sso = importlib.import_module("NFACT.decomp.decomposition.sso.nmfsso")
    ref_nmf_decomp = ref_sso.nmf_decomp
    assert ref_nmf_decomp is not nfact_b200.nmf_decomp
    # calling conventions first, against the reference's own functions: the reference's parameters, in order,
    # are a prefix of ours (ours may add keyword-only / defaulted extras)
    checked = 0
    for modname, names in dropin._DEFINITIONS.items():
        try:
            module = importlib.import_module(modname)
        except Exception:
            continue
        for name in names:
            ref_fn, ours = getattr(module, name), dropin._replacement(name)
            if inspect.isclass(ref_fn):
                ref_fn, ours = ref_fn.__init__, ours.__init__
            ref_params = list(inspect.signature(ref_fn).parameters.values())
            our_params = list(inspect.signature(ours).parameters.values())
            assert [p.name for p in our_params[:len(ref_params)]] == [p.name for p in ref_params], (modname, name)
            for rp, op in zip(ref_params, our_params):
                assert (rp.default is inspect.Parameter.empty) == (op.default is inspect.Parameter.empty), (name, rp.name)
            for extra in our_params[len(ref_params):]:
                assert extra.default is not inspect.Parameter.empty, (name, extra.name)
            checked += 1
    assert checked >= 12

    patched = dropin.install()
    assert "NFACT.decomp.decomposition.sso.nmfsso.nmf_decomp" in patched
    assert ref_sso.nmf_decomp is nfact_b200.nmf_decomp
    assert ref_sso.thresholding is nfact_b200.thresholding                 # a from-import copy, rebound too
    md = importlib.import_module("NFACT.decomp.decomposition.matrix_decomposition")
    assert md.nmf_sso is nfact_b200.nmf_sso and md.get_parameters is dropin.get_parameters
    # NMF: nfact_b200's mirror of the defaults; ICA (out of scope) goes through to the reference's own function
    assert md.get_parameters(None, "nmf", 7) == nfact_b200.get_parameters(None, "nmf", 7)
    ica = md.get_parameters(None, "ica", 7)
    ref_get = dropin._originals[("NFACT.decomp.decomposition.matrix_decomposition", "get_parameters")]
    assert ica == ref_get(None, "ica", 7) and ica["n_components"] == 7 and "alpha_W" not in ica
    mh = importlib.import_module("NFACT.decomp.decomposition.matrix_handling")
    assert mh.process_fdt_matrix2 is nfact_b200.process_fdt_matrix2 and mh.avg_fdt is nfact_b200.avg_fdt
    drm = importlib.import_module("NFACT.dual_reg.dual_regression.dual_regression_methods")
    assert drm.nmf_dual_regression is nfact_b200.nmf_dual_regression and drm.run_decomp is nfact_b200.run_decomp
    assert drm.ica_dual_regression.__module__.startswith("NFACT.")          # ICA is out of scope: untouched
    assert dropin.install() == []                                            # idempotent
    assert set(dropin.installed()) == set(patched)

    dropin.uninstall()
    assert ref_sso.nmf_decomp is ref_nmf_decomp and dropin.installed() == []


def test_entry_point_modules_pick_up_the_rebound_functions(reference):
    """`python -m nfact_b200.dropin nfact_decomp ...`: the mains `from`-import the seams after install()."""
    dropin.install()
    bound = 0
    for tool, (modname, func) in dropin._ENTRY_POINTS.items():
        try:
            module = importlib.import_module(modname)
        except Exception:
            continue                       # a dependency of the reference's image writers that cannot be stubbed
        assert callable(getattr(module, func))
        dropin.install()
        for name in dropin._IMPORTERS.get(modname, ()):
            assert getattr(module, name) is dropin._replacement(name), (modname, name)
        bound += 1
    run_dr = sys.modules.get("NFACT.dual_reg.dual_regression.run_dual_regression")
    if run_dr is not None:
        assert run_dr.nmf_dual_regression is nfact_b200.nmf_dual_regression
        assert run_dr.load_fdt_matrix is nfact_b200.load_fdt_matrix
        bound += 1
    assert bound >= 1


def test_usage_message():
    with pytest.raises(SystemExit):
        dropin.main(["nfact_pp"])


def test_nfact_dr_subject_loop_is_rebound_and_keeps_the_reference_calls(reference, monkeypatch, tmp_path):
    """run_locally(args, paths) of set_up_dual_regression.py:186-235: after install() nfact_dr's main calls the
    nfact_b200 loop, which fetches the group components and saves every subject through the reference's own
    functions with the reference's arguments (checked against their real signatures); ICA keeps the reference loop."""
    try:
        setup = importlib.import_module("NFACT.dual_reg.dual_regression.set_up_dual_regression")
        drf = importlib.import_module("NFACT.dual_reg.nfact_dr_functions")
    except Exception as exc:
        pytest.skip(f"reference dual-regression modules need an image library that cannot be stubbed: {exc}")
    reference_loop = setup.run_locally
    dropin.install()
    assert setup.run_locally is dropin.run_locally and reference_loop is not dropin.run_locally

    subjects = [str(tmp_path / "subjects" / f"sub_{i:02d}" / "omatrix2") for i in range(3)]
    args = {"algo": "nmf", "ptxdir": subjects, "seeds": ["l.surf.gii", "r.surf.gii"], "roi": ["l_roi.gii", "r_roi.gii"],
            "outdir": str(tmp_path / "out"), "threshold": "3", "normalise": True, "cifti": False, "n_cores": 2}
    paths = {"component_path": str(tmp_path / "comp"), "group_average_path": str(tmp_path / "avg")}
    calls = {"components": [], "saved": [], "loop": []}
    save_sig = inspect.signature(drf.save_dual_regression_images)
    comp_sig = inspect.signature(drf.get_group_level_components)

    def fake_components(*a, **k):
        bound = comp_sig.bind(*a, **k)
        calls["components"].append(tuple(bound.arguments.values()))
        return {"grey_components": np.ones((4, 2)), "white_components": np.ones((2, 5))}

    def fake_save(*a, **k):
        bound = save_sig.bind(*a, **k)
        bound.apply_defaults()
        calls["saved"].append(bound.arguments)

    def fake_loop(list_of_subject_dirs, components, seeds, threshold=3, normalise=False, rank=0, world=1, save=None,
                  prefetch=True, read_ahead=0):
        calls["loop"].append((list(list_of_subject_dirs), seeds, threshold, normalise))
        for d in list_of_subject_dirs:
            save(d, {"grey_components": np.zeros((4, 2)), "white_components": np.zeros((2, 5)),
                     "normalised_grey": np.zeros((4, 2)), "normalised_white": np.zeros((2, 5))})
        return {d: None for d in list_of_subject_dirs}

    monkeypatch.setattr(drf, "get_group_level_components", fake_components)
    monkeypatch.setattr(drf, "save_dual_regression_images", fake_save)
    monkeypatch.setattr(nfact_b200, "run_locally", fake_loop)
    assert setup.run_locally(args, paths) is None
    assert calls["components"] == [(paths["component_path"], paths["group_average_path"], args["seeds"], args["roi"])]
    assert calls["loop"] == [(subjects, args["seeds"], 3, True)]
    assert [c["ptx_directory"] for c in calls["saved"]] == subjects
    first = calls["saved"][0]
    assert first["nfact_path"] == os.path.join(args["outdir"], "nfact_dr") and first["algo"] == "NMF"
    assert first["dim"] == 2 and first["seeds"] == args["seeds"] and first["roi"] == args["roi"]
    assert first["sub"] == drf.get_subject_id(subjects[0]) == "sub_00" and first["cifti_save"] is False
    assert "normalised_white" in first["components"]

    seen = []
    monkeypatch.setitem(dropin._originals, dropin._DR_LOOP, lambda a, p: seen.append(a["algo"]))
    dropin.run_locally(dict(args, algo="ica"), paths)
    assert seen == ["ica"] and len(calls["loop"]) == 1           # ICA went to the reference's loop
