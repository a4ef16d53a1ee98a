"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (mtn_b200/).

Loads the UNMODIFIED reference implementation from ``/root/reference`` (or
``$MTN_REF_DIR``) so that the restatement in ``oracle/mtn_oracle.py`` can be
pinned against the real thing and golden fixtures can be generated
(``oracle/make_golden.py``).  The reference cannot travel to the GPU box, so
nothing under ``tests/ -m gpu``, ``bench.py`` or ``__graft_entry__.smoke()``
calls into this module.

Two shims are needed, neither touches the reference:

* ``data_utils.py:8`` does ``from torchtext import data, datasets`` and
  ``data_utils.py:69`` subclasses ``data.Iterator`` at import time.  torchtext is
  not installed; stub modules are pre-registered in ``sys.modules``.
* ``data_utils.py:28`` hard-codes ``.cuda()`` inside ``Batch.__init__`` for the
  feature tensors.  ``make_cpu_batch`` builds the Batch with ``fts=None`` and
  then fills ``fts`` / ``fts_mask`` with exactly the arithmetic of
  ``data_utils.py:29-30``.
"""
import os
import sys
import types
import warnings

import torch

_REF = None


_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ref_dir(allow_build_container_path=True):
    """SURVEY 8c lookup order: $MTN_REF_DIR -> /root/reference (the build container only; bench.py passes
    allow_build_container_path=False because nothing it runs on the GPU box may read that path) -> baseline/_ref/
    (a driver-installed copy that travels with the repo snapshot)."""
    cands = [os.environ.get("MTN_REF_DIR")]
    if allow_build_container_path:
        cands.append("/root/reference")
    cands.append(os.path.join(_REPO, "baseline", "_ref"))
    for cand in cands:
        if cand and os.path.isfile(os.path.join(cand, "mtn.py")):
            return cand
    return None


def available():
    return ref_dir() is not None


def load(d=None):
    """Return (mtn, data_utils) modules of the reference (from `d`, default: ref_dir())."""
    global _REF
    if _REF is not None:
        return _REF
    d = d or ref_dir()
    if d is None:
        raise RuntimeError("reference checkout not found (set MTN_REF_DIR)")
    if "torchtext" not in sys.modules:
        tt = types.ModuleType("torchtext")
        tt_data = types.ModuleType("torchtext.data")
        tt_ds = types.ModuleType("torchtext.datasets")

        class Iterator(object):
            pass

        tt_data.Iterator = Iterator
        tt_data.batch = lambda *a, **k: iter(())
        tt.data, tt.datasets = tt_data, tt_ds
        sys.modules.update({"torchtext": tt, "torchtext.data": tt_data,
                            "torchtext.datasets": tt_ds})
    # The reference's module names (mtn, data_utils) are imported under private
    # aliases so they can never shadow the product package.
    saved = {k: sys.modules.get(k) for k in ("mtn", "data_utils")}
    sys.path.insert(0, d)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for k in ("mtn", "data_utils"):
                sys.modules.pop(k, None)
            import data_utils as ref_du  # noqa
            import mtn as ref_mtn  # noqa
    finally:
        sys.path.remove(d)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)
    _REF = (ref_mtn, ref_du)
    return _REF


def make_cpu_batch(query, his, cap, trg, trg_y, fts, pad=1):
    """Reference ``Batch`` on CPU.  ``fts`` is a list of (B, L, F) float tensors
    (already batch-major; the reference permutes from (L, B, F))."""
    _, du = load()
    b = du.Batch(query, his, None, None, cap, trg, trg_y, pad)
    if fts is not None:
        # data_utils.py:29-30, minus the .cuda()/.permute()
        b.fts_mask = [(torch.sum(ft != 1, dim=2) != 0).unsqueeze(-2) for ft in fts]
        b.fts = [ft * b.fts_mask[i].squeeze(1).unsqueeze(-1).expand_as(ft).float()
                 for i, ft in enumerate(fts)]
    return b
