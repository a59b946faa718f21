"""Import the UNMODIFIED reference (ai-forever/CerberusDet) in this container.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/gen_golden.py`` to produce the golden
vectors under ``tests/golden/`` and by the container-only tests that validate the
oracle restatement against the real reference.  ``/root/reference`` does not exist
on the GPU box, so nothing in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may
call this module.

The reference's hot-path modules import plotting/logging packages at module import
time (``utils/general.py:20`` -> ``utils/metrics.py:9`` -> matplotlib;
``models/yolo.py:29`` -> ``utils/plots.py:12-16`` -> seaborn;
``models/experimental.py:8`` -> ``utils/mlflow_logging.py:7-11`` -> mlflow).  None of
them is used by decode or NMS, so inert placeholder modules are registered first.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CERB_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "cerberusdet"))


class _Inert(types.ModuleType):
    """A module whose every attribute is a do-nothing callable / sub-object."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        def _noop(*a, **k):
            return None

        return _noop


def _install_placeholders() -> None:
    names = [
        "matplotlib",
        "matplotlib.pyplot",
        "seaborn",
        "mlflow",
        "mlflow.models",
        "mlflow.models.signature",
        "mlflow.tracking",
    ]
    for n in names:
        try:
            __import__(n)
            continue
        except Exception:
            pass
        mod = _Inert(n)
        mod.__path__ = []  # behave like a package so "import a.b" resolves
        sys.modules[n] = mod
        if "." in n:
            parent, child = n.rsplit(".", 1)
            setattr(sys.modules[parent], child, mod)


_loaded = None


def load_reference():
    """Return a namespace with the reference's hot-path symbols (imported verbatim)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the tree is read-only
    _install_placeholders()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import cerberusdet.utils.general as general  # noqa: E402
    import cerberusdet.utils.tal as tal  # noqa: E402
    import cerberusdet.models.yolo as yolo  # noqa: E402

    ns = types.SimpleNamespace(general=general, tal=tal, yolo=yolo)
    _loaded = ns
    return ns
