"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference (`/root/reference/hwer`) in this container.

The reference cannot be imported as shipped here (`import hwer` pulls in DGL 0.4, bidict, more_itertools,
matplotlib, ... none of which are installed).  This shim makes the *hot-path* modules importable without
touching the reference tree:

  * a stub package object `hwer` whose `__path__` points at the reference, so `hwer/__init__.py` (which
    imports everything) is skipped;
  * a dict-backed `bidict` stand-in exposing `.inverse` (used at recommendation_base.py:73,81,89,98,101-102,147);
  * inert stub modules for heavyweight third-party imports that the serving path never calls.

Only `oracle/make_golden.py` and tests that are skipped when `/root/reference` is absent use it; nothing
in the product imports this file.  The GPU box has no `/root/reference`, so nothing run there may need it.
"""
import importlib
import importlib.abc
import importlib.machinery
import itertools
import os
import random
import sys
import types

REFERENCE_ROOT = os.environ.get("HWER_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "hwer"))


class bidict(dict):
    """Minimal stand-in for bidict==0.18.3: a dict plus a lazily rebuilt inverse view."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._inv = None

    @property
    def inverse(self):
        if self._inv is None or len(self._inv) != len(self):
            self._inv = {v: k for k, v in self.items()}
        return self._inv

    def __setitem__(self, k, v):
        self._inv = None
        super().__setitem__(k, v)

    def update(self, *a, **k):
        self._inv = None
        super().update(*a, **k)


class _Anything:
    """Attribute/call sink for stubbed third-party modules."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    __all__ = []
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


_STUB_ROOTS = ("dgl", "more_itertools", "dill", "matplotlib", "seaborn", "torch_optimizer", "tensorflow", "flair",
               "fasttext", "hnswlib", "nmslib", "surprise", "hyperopt", "gensim", "nltk", "bs4", "contractions",
               "unidecode", "stanfordnlp", "swifter", "MulticoreTSNE", "umap")


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        if spec.name == "more_itertools":
            m.flatten = itertools.chain.from_iterable

            def chunked(it, n):
                it = iter(it)
                while True:
                    c = list(itertools.islice(it, n))
                    if not c:
                        return
                    yield c
            m.chunked = chunked
        return m

    def exec_module(self, module):
        pass


_loaded = {}


def load_reference():
    """Returns a namespace with the reference's hot-path modules (recommendation_base, utils, validation, gcn_ncf)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not reference_available():
        raise RuntimeError("reference checkout not found at %s" % REFERENCE_ROOT)
    bd = types.ModuleType("bidict")
    bd.bidict = bidict
    sys.modules.setdefault("bidict", bd)
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())   # appended: real installed modules always win
    pkg = types.ModuleType("hwer")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "hwer")]
    sys.modules["hwer"] = pkg
    rb = importlib.import_module("hwer.recommendation_base")
    ut = importlib.import_module("hwer.utils")
    _loaded.update(recommendation_base=rb, utils=ut)
    # validation.py:80 calls random.sample on a set, removed in Python 3.11
    _orig_sample = random.sample

    def _sample(population, k, **kw):
        if isinstance(population, (set, frozenset)):
            population = sorted(population, key=repr)
        return _orig_sample(population, k, **kw)
    random.sample = _sample
    try:
        _loaded["validation"] = importlib.import_module("hwer.validation")
    except Exception as e:  # pragma: no cover - informational
        _loaded["validation"] = None
        _loaded["validation_error"] = repr(e)
    try:
        _loaded["gcn_ncf"] = importlib.import_module("hwer.gcn_ncf")
    except Exception as e:  # pragma: no cover
        _loaded["gcn_ncf"] = None
        _loaded["gcn_ncf_error"] = repr(e)
    return types.SimpleNamespace(**_loaded)


if __name__ == "__main__":
    ns = load_reference()
    print({k: (v if isinstance(v, str) else bool(v)) for k, v in vars(ns).items()})
