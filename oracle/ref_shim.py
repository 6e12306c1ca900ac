"""ORACLE helper (container-only): import the UNMODIFIED reference modules from /root/reference with
the third-party modules that are absent here stubbed out.  Used by
tests/golden/make_golden.py to generate fixtures and by tests that are skipped when the reference
tree is not present (it does not exist on the GPU box)."""
import os
import sys
import types

REF = os.environ.get("CLSKD_REFERENCE", "/root/reference")
_LOCAL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")     # oracle/make_ref.py's unmodified copy


def _root():
    for r in (REF, _LOCAL):
        if os.path.isdir(r) and os.path.exists(os.path.join(r, "DCCRN.py")):
            return r
    return None


def available():
    return _root() is not None


def apply_torch_compat():
    """torch-version shims the unmodified reference needs on torch >= 2 / a CPU box (documented in
    tests/golden/make_golden.py): torch.stft without return_complex (framework.py:27) and the hard-coded
    .cuda() calls of ABF.__init__ (framework.py:198-202) when no GPU is used."""
    import torch
    if getattr(torch.stft, "_clskd_compat", False):
        return
    _stft = torch.stft

    def stft_compat(*a, **k):
        k.setdefault("return_complex", True)
        out = _stft(*a, **k)
        return torch.view_as_real(out) if out.is_complex() else out
    stft_compat._clskd_compat = True
    torch.stft = stft_compat
    torch.nn.Module.cuda = lambda self, device=None: self
    torch.Tensor.cuda = lambda self, *a, **k: self


def load():
    """-> dict of the reference's modules (DCCRN, config, tools_for_model, tools_for_loss, framework,
    feature_extraction), imported under their own top-level names."""
    root = _root()
    if root is None:
        raise RuntimeError("reference tree not found at %s (nor oracle/_ref)" % REF)

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
    if root not in sys.path:
        sys.path.insert(0, root)
    import importlib
    mods = {}
    for n in ('config', 'tools_for_model', 'tools_for_loss', 'DCCRN', 'feature_extraction', 'framework'):
        mods[n] = importlib.import_module(n)
    return mods
