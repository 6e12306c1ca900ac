"""ORACLE helper (container-only): import the UNMODIFIED reference modules from /root/reference with
the third-party modules that are absent here stubbed out.  Used by
tests/golden/make_golden.py to generate fixtures and by tests that are skipped when the reference
tree is not present (it does not exist on the GPU box)."""
import os
import sys
import types

REF = os.environ.get("CLSKD_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(REF) and os.path.exists(os.path.join(REF, "DCCRN.py"))


def load():
    """-> dict of the reference's modules (DCCRN, config, tools_for_model, tools_for_loss, framework,
    feature_extraction), imported under their own top-level names."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF)

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    class _Ctor:            # tools_for_loss.py:258-259 instantiates these at import time
        def __init__(self, *a, **k):
            pass

    mpl = stub('matplotlib')
    mpl.pylab = stub('matplotlib.pylab')
    stub('pesq', pesq=None)
    stub('pystoi', stoi=None)
    stub('asteroid')
    stub('asteroid.losses', SingleSrcPMSQE=_Ctor, PITLossWrapper=_Ctor)
    stub('asteroid_filterbanks', STFTFB=_Ctor, Encoder=_Ctor, transforms=None)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import importlib
    mods = {}
    for n in ('config', 'tools_for_model', 'tools_for_loss', 'DCCRN', 'feature_extraction', 'framework'):
        mods[n] = importlib.import_module(n)
    return mods
