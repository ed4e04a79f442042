"""Drop-in aliasing: makes `import procyon.model.model_unified` (etc.) resolve to procyon_b200.

The reference's drivers (scripts/*.py, procyon/evaluate/framework/procyon.py, procyon/inference/retrieval_utils.py)
import the hot path by the dotted names below, and checkpoints pickle `procyon.training.training_args_IT.*`
instances. `install()` registers aliases for exactly those modules; it never shadows a real `procyon` package that
is already imported (in that case use `patch_reference()` from INTEGRATION.md instead).
"""
from __future__ import annotations

import importlib
import sys
import types

_ALIASES = {
    "procyon.model.model_unified": "procyon_b200.model.model_unified",
    "procyon.model.esm": "procyon_b200.model.esm",
    "procyon.model.pmc_llama": "procyon_b200.model.pmc_llama",
    "procyon.model.model_utils": "procyon_b200.model.model_utils",
    "procyon.model.contrastive": "procyon_b200.model.contrastive",
    "procyon.training.training_args_IT": "procyon_b200.training.training_args_IT",
    "procyon.training.train_utils": "procyon_b200.training.train_utils",
    "procyon.inference.retrieval_utils": "procyon_b200.inference.retrieval_utils",
    "procyon.data.inference_utils": "procyon_b200.data.inference_utils",
    "procyon.evaluate.framework.procyon": "procyon_b200.evaluate.framework.procyon",
}


def install(force: bool = False) -> bool:
    real = sys.modules.get("procyon")
    if real is not None and getattr(real, "__file__", None) and not force:
        return False  # the real package is loaded; do not shadow it
    for pkg in ("procyon", "procyon.model", "procyon.training", "procyon.inference", "procyon.data",
                "procyon.evaluate", "procyon.evaluate.framework"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
    for alias, target in _ALIASES.items():
        try:
            mod = importlib.import_module(target)
        except ModuleNotFoundError:
            continue
        sys.modules[alias] = mod
        parent, _, leaf = alias.rpartition(".")
        setattr(sys.modules[parent], leaf, mod)
    return True


def patch_reference():
    """With the real reference package importable, swap its hot-path classes for the B200 ones in place."""
    import procyon.model.model_unified as ref_mu  # noqa: F401  (the real one)

    from .model import esm, model_unified, model_utils, pmc_llama

    ref_mu.UnifiedProCyon = model_unified.UnifiedProCyon
    ref_mu.ESM_PLM = esm.ESM_PLM
    ref_mu.LlamaPostTokenization = pmc_llama.LlamaPostTokenization
    ref_mu.create_mlp = model_utils.create_mlp
    try:  # the evaluation plugins core.py registers under "ProCyon" (evaluate/framework/core.py:68-110)
        import procyon.evaluate.framework.core as ref_core

        from .evaluate.framework import procyon as plugins

        ref_core.caption_models["ProCyon"] = plugins.ProcyonCaptionEval
        ref_core.qa_models["ProCyon"] = plugins.ProcyonQAEval
        ref_core.retrieval_models["ProCyon"] = plugins.ProcyonRetrievalEval
    except ImportError:  # the reference's evaluate package has heavy optional dependencies
        pass
