"""TEST INFRASTRUCTURE ONLY -- import the unmodified reference in THIS container.

``/root/reference`` exists only in the build container (never on the GPU box), so this
module is used by exactly one thing: ``tests/golden/make_golden.py``, which runs the
reference on seeded inputs and commits the outputs as fixtures.  It must never be
imported by the product package, by ``bench.py`` or by ``-m gpu`` tests.

Three modules the reference imports at module scope are absent here and are stubbed:
``matplotlib`` (util.py:8), ``webdataset`` (dataset.py:3) and ``torch_dct``
(util.py:9, replaced by ``oracle/torch_dct_standin.py``).
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "dct_autoencoder"))


def import_reference():
    """Returns the imported ``dct_autoencoder`` package of the reference."""
    if not reference_available():
        raise RuntimeError("the reference tree is only present in the build container")
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    import torch_dct_standin

    sys.modules.setdefault("torch_dct", torch_dct_standin)
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if "webdataset" not in sys.modules:
        wds = types.ModuleType("webdataset")
        wds.WebDataset = object
        wds.handlers = types.SimpleNamespace(warn_and_continue=None)
        sys.modules["webdataset"] = wds
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import dct_autoencoder  # noqa: E402  (the reference package)

    return dct_autoencoder
