"""Registers `transformers.models.markushgrapher` so the reference's own import
(markushgrapher/core/common/begin.py:7-13) resolves to this package:

    import markushgrapher_b200.hf_shim  # once, before `import markushgrapher.core...`
"""
import sys
import types

from .configuration import MarkushgrapherConfig
from .modeling import MarkushgrapherForConditionalGeneration
from .processing import MarkushgrapherImageProcessor, MarkushgrapherProcessor, MarkushgrapherTokenizer

EXPORTS = {
    "MarkushgrapherConfig": MarkushgrapherConfig,
    "MarkushgrapherForConditionalGeneration": MarkushgrapherForConditionalGeneration,
    "MarkushgrapherImageProcessor": MarkushgrapherImageProcessor,
    "MarkushgrapherProcessor": MarkushgrapherProcessor,
    "MarkushgrapherTokenizer": MarkushgrapherTokenizer,
}


def install() -> types.ModuleType:
    import transformers.models as tm

    mod = types.ModuleType("transformers.models.markushgrapher")
    mod.__dict__.update(EXPORTS)
    mod.__all__ = list(EXPORTS)
    sys.modules["transformers.models.markushgrapher"] = mod
    setattr(tm, "markushgrapher", mod)
    return mod


install()
