"""Test helper: import the REFERENCE's own Python modules (from /root/reference, never copied) in this container.

The reference's glue imports third-party packages that are absent here (no network): matplotlib, albumentations,
rdkit, SmilesPE, nltk, rouge_score, markushgenerator, molscribe, clearml.  None of them is on the path the tests
exercise, so a meta-path finder hands out empty stand-in modules for exactly those names.  Two environment
adaptations besides that, neither touching the reference's logic:
  * torch._utils._accumulate (a private symbol the reference's utils.py imports, gone in torch 2.11) = itertools.accumulate;
  * transformers.TrainingArguments._setup_devices needs `accelerate` (absent): answered with the CPU device.
Only tests / golden generators use this; it needs /root/reference and therefore only runs in the build container."""
import functools
import importlib.abc
import importlib.machinery
import itertools
import os
import sys
import types

REF = "/root/reference"
ABSENT = ("matplotlib", "albumentations", "rdkit", "SmilesPE", "nltk", "rouge_score", "markushgenerator", "molscribe",
          "clearml")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "markushgrapher"))


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, n):
        return _Anything()


class _Stub(types.ModuleType):
    __path__ = []

    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Anything


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in ABSENT:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        return _Stub(spec.name)

    def exec_module(self, module):
        pass


_installed = False


def install():
    """make `import markushgrapher...` (the reference package) work in this process, with the repo's
    transformers.models.markushgrapher shim registered first (reference begin.py:7-13 imports it)"""
    global _installed
    if _installed:
        return
    import torch
    import torch._utils

    if not hasattr(torch._utils, "_accumulate"):
        torch._utils._accumulate = itertools.accumulate
    import markushgrapher_b200.hf_shim  # noqa: F401
    import transformers.training_args as ta

    p = functools.cached_property(lambda self: torch.device("cpu"))
    p.__set_name__(ta.TrainingArguments, "_setup_devices")
    ta.TrainingArguments._setup_devices = p
    sys.meta_path.insert(0, _Finder())
    if REF not in sys.path:
        sys.path.insert(0, REF)
    _installed = True
