"""Import the UNMODIFIED reference (ai-forever/CerberusDet) in this container.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/gen_golden.py`` to produce the golden
vectors under ``tests/golden/`` and by the tests that validate the oracle restatement and
the drop-in against the real reference.  ``/root/reference`` does not exist on the GPU box;
there the same files are found under ``oracle/_ref`` (``oracle/make_ref.py``, git-ignored, shipped
with the snapshot), so the ``-m gpu`` drop-in tests and ``bench.py``'s reference legs run the
reference itself and never read ``/root/reference``.

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

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    """``/root/reference`` in the build container; on the GPU box the file-for-file copy ``oracle/_ref`` that
    ``oracle/make_ref.py`` ships with the snapshot (git-ignored)."""
    env = os.environ.get("CERB_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/cerberusdet"):
        return "/root/reference"
    return os.path.join(_HERE, "_ref")


REFERENCE_ROOT = _find_root()


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
