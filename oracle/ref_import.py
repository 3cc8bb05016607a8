"""Import the UNMODIFIED reference (`/root/reference`) on CPU, for golden-vector generation only.

TEST INFRASTRUCTURE.  This file only works in the build container, where `/root/reference` is
mounted; nothing that runs on the GPU box may import it.  It is used by
`tests/golden/make_golden.py` (fixture generation) and by `tests/test_oracle_vs_reference.py`
(which skips itself when the reference tree is absent).

The reference needs a number of third-party modules at import time that are absent here
(SURVEY.md section 8c): tensorflow (model/__init__.py:4 -> deepspeaker), matplotlib
(utils/tools.py:9-14), librosa/parselmouth/pyworld/pycwt (utils/pitch_tools.py:4-10),
unidecode / inflect (text/), python_speech_features (deepspeaker/audio_ds.py).  None of them is
touched by the acoustic-model forward path, so empty stub modules are enough.
"""
import importlib.machinery
import os
import sys
import types

REF_ROOT = os.environ.get("CTTS_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "tensorflow", "tensorflow.keras", "tensorflow.keras.backend", "tensorflow.keras.layers",
    "tensorflow.keras.models", "tensorflow.keras.optimizers", "tensorflow.keras.regularizers",
    "tensorflow.keras.callbacks", "tensorflow.keras.utils",
    "matplotlib", "matplotlib.pyplot", "librosa", "librosa.filters", "librosa.util", "parselmouth", "pyworld",
    "pycwt", "unidecode", "inflect", "python_speech_features", "tgt", "g2p_en", "pypinyin",
    "pyloudnorm",
]


class _Anything:
    """Attribute sink: any attribute / call / subclassing on a stub returns another sink."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __mro_entries__(self, bases):
        return (object,)


def _make_stub(name):
    mod = types.ModuleType(name)
    mod.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    mod.__path__ = []  # behave like a package so that `import a.b` works

    def _getattr(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Anything()

    mod.__getattr__ = _getattr  # type: ignore[attr-defined]
    return mod


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "model"))


def import_reference():
    """Returns the reference's `model` package and `utils.tools` module, imported from REF_ROOT.

    Side effects: chdir(REF_ROOT) (the reference opens ./config and ./preprocessed_data with
    relative paths: utils/tools.py:20, config/LJSpeech/preprocess.yaml:7) and sys.path[0]=REF_ROOT.
    """
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    import torch  # noqa: F401  (import the real heavy modules before any stub exists)
    import numba  # noqa: F401
    import scipy.io, scipy.interpolate, sklearn.manifold  # noqa: F401,E401
    for name in _STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _make_stub(name)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    os.chdir(REF_ROOT)
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int  # utils/pitch_tools.py:34 uses the removed alias (numpy branch only)
    import model as ref_model  # noqa
    import utils.tools as ref_tools  # noqa
    return ref_model, ref_tools


def reference_configs(dataset="LJSpeech"):
    import numpy as np
    _, ref_tools = import_reference()
    p, m, t = ref_tools.get_configs_of(dataset)
    # train.py:229-231 / synthesize.py:177-179 patch the CWT scales in at run time; only len() is used
    p["preprocessing"]["pitch"]["cwt_scales"] = 0.01 * 2.0 ** np.arange(10)
    return p, m, t
