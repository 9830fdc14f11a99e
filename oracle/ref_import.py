"""
Import shim for the *reference* GPry package (test infrastructure, container-only).

The reference at ``/root/reference`` imports ``matplotlib`` and ``getdist`` at module import
time (gpry/plots.py:12-14,29 and gpry/mc.py:12-13); neither is installed and neither is on
the GP hot path.  This module injects empty stand-ins for them so that ``import gpry`` works,
and is used ONLY by ``oracle/gen_golden.py`` (golden-vector generation) and by the optional
``tests/test_oracle_vs_reference.py`` cross-check, both of which run in the build container
where ``/root/reference`` exists.  Nothing that runs on the GPU box imports this file.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GPRY_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gpry"))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def import_reference():
    """Returns the imported reference ``gpry`` package (read-only tree: no bytecode)."""
    if "gpry" in sys.modules and getattr(sys.modules["gpry"], "__file__", "").startswith(
            REFERENCE_ROOT):
        return sys.modules["gpry"]
    if not reference_available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True

    class _Anything:  # accepts any construction / attribute access
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, name):
            return _Anything()

        def __call__(self, *a, **k):
            return _Anything()

    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib", rcParams={}, use=lambda *a, **k: None)
        mpl.pyplot = _stub("matplotlib.pyplot", rcParams={})
        mpl.cm = _stub("matplotlib.cm")
        mpl.colors = _stub("matplotlib.colors")
        mpl.ticker = _stub("matplotlib.ticker", MaxNLocator=_Anything)
        mpl.lines = _stub("matplotlib.lines", Line2D=_Anything)
        mpl.patches = _stub("matplotlib.patches", Patch=_Anything)
    if "getdist" not in sys.modules:
        gd = _stub("getdist", MCSamples=_Anything)
        gd.mcsamples = _stub("getdist.mcsamples", MCSamples=_Anything,
                             loadMCSamples=_Anything())
        gd.gaussian_mixtures = _stub("getdist.gaussian_mixtures", GaussianND=_Anything)
        gd.plots = _stub("getdist.plots")
        gd.densities = _stub("getdist.densities")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import gpry  # noqa
    return gpry
