"""Import the UNMODIFIED reference functions (only where /root/reference exists: the build container).

TEST INFRASTRUCTURE ONLY.  Used by tests/test_oracle_vs_reference.py and oracle/make_golden.py to pin the oracle and
to generate the committed golden fixtures.  Nothing at GPU run time reads /root/reference.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import warnings

REF_ROOT = os.environ.get("RIR_REFERENCE_ROOT", "/root/reference")
REF_BENCH = os.path.join(REF_ROOT, "src", "benchmark")
# The unmodified reference also installs offline (pip --no-deps --target baseline/_ref); that copy is
# git-ignored but travels to the GPU box, so the oracle can be checked against the real reference there too.
_INSTALLED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "benchmark")
if not os.path.isfile(os.path.join(REF_BENCH, "utils", "evaluate.py")) and os.path.isfile(
        os.path.join(_INSTALLED, "utils", "evaluate.py")):
    REF_ROOT, REF_BENCH = os.path.dirname(_INSTALLED), _INSTALLED


def available() -> bool:
    return os.path.isfile(os.path.join(REF_BENCH, "utils", "evaluate.py"))


@contextlib.contextmanager
def _ref_path():
    sys.path.insert(0, REF_BENCH)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            yield
    finally:
        sys.path.remove(REF_BENCH)


def load():
    """Returns a namespace of reference callables.  Raises RuntimeError when the reference tree is absent."""
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    # the reference's top-level package names (`utils`, `networks`, `models`) are generic: import them under the
    # reference path, then drop them from sys.modules so they cannot shadow anything else.
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in ("utils", "networks", "models", "config")}
    with _ref_path():
        import importlib

        ev = importlib.import_module("utils.evaluate")
        hf = importlib.import_module("utils.helpfunc")
        rn = importlib.import_module("networks.RetrievalNet")
        bb = importlib.import_module("networks.backbone")
        sp = importlib.import_module("networks.spca")
        gp = importlib.import_module("models.gem_pooling")
        g2 = importlib.import_module("models.senet_g2")
        spp = importlib.import_module("models.spoc")
    ns = type("Reference", (), {})()
    ns.compute_ap, ns.compute_map, ns.compute_map_and_print = ev.compute_ap, ev.compute_map, ev.compute_map_and_print
    ns.extract_vectors = hf.extract_vectors
    ns.gem, ns.spoc = rn.gem, rn.spoc
    ns.pcawhitenlearn_shrinkage = bb.pcawhitenlearn_shrinkage
    ns.ConvDimReduction = sp.ConvDimReduction
    ns.GeMPooling, ns.G2Pooling, ns.SpatialPyramidPooling = gp.GeMPooling, g2.G2Pooling, spp.SpatialPyramidPooling
    for k in list(sys.modules):
        if k.split(".")[0] in ("utils", "networks", "models", "config"):
            sys.modules.pop(k)
    sys.modules.update(saved)
    return ns


def load_iris_copy():
    """The DUPLICATE evaluation functions inside the reference's entry script (iris_evaluate.py:11-265).  The script
    itself cannot be imported (it imports the missing `iris_implementation`, :9), so the three function definitions are
    cut out of the unmodified file by line range (ast) and executed as they are."""
    import ast

    import numpy as np
    path = os.path.join(REF_BENCH, "iris_evaluate.py")
    if not os.path.isfile(path):
        raise RuntimeError(f"{path} not found")
    src = open(path, encoding="utf-8").read()
    tree = ast.parse(src)
    lines = src.splitlines(keepends=True)
    ns = {"np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("compute_ap", "compute_map", "compute_map_and_print"):
            exec(compile("".join(lines[node.lineno - 1:node.end_lineno]), path, "exec"), ns)
    out = type("IrisCopy", (), {})()
    out.compute_ap, out.compute_map, out.compute_map_and_print = ns["compute_ap"], ns["compute_map"], ns["compute_map_and_print"]
    return out


def quiet(fn, *a, **kw):
    """Call a reference function that prints, returning (result, printed_text)."""
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = fn(*a, **kw)
    return out, buf.getvalue()
