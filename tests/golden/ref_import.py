"""Import pieces of the UNMODIFIED reference (/root/reference) in this container, for golden-vector generation.

The reference's hot-path modules import third-party packages that are absent here (deepspeed, fair-esm, peft,
evaluate, ...). None of them is needed by the host-level functions we execute, so a meta-path hook serves inert
stub modules for those top-level names. Nothing from the reference is copied: its functions run from where they lie.
Only used by make_golden.py (never at test time; /root/reference does not exist on the GPU box).
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"
STUB_TOP = {"deepspeed", "esm", "peft", "evaluate", "bitsandbytes", "wandb", "pynvml", "captum", "Bio",
            "torch_geometric", "torchdrug", "dotenv", "bert_score", "rouge_score", "nltk", "loguru", "accelerate",
            "flash_attn", "torch_scatter", "rdkit", "fastapi", "uvicorn"}


class _DummyMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy


class _Dummy(metaclass=_DummyMeta):
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Dummy()

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in STUB_TOP:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        if module.__name__ == "dotenv":
            module.load_dotenv = lambda *a, **k: None


def install():
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError("/root/reference is not available (golden vectors are generated in the build container)")
    os.environ.setdefault("DATA_DIR", "/tmp/procyon_data")
    os.environ.setdefault("HOME_DIR", "/tmp/procyon_home")
    os.environ.setdefault("LLAMA3_PATH", "/tmp/llama3")
    # let transformers probe its optional dependencies BEFORE the stubs exist (it would mistake them for installs)
    import transformers  # noqa: F401
    import transformers.models.esm.modeling_esm  # noqa: F401
    import transformers.models.llama.modeling_llama  # noqa: F401
    from transformers import (AutoModelForMaskedLM, AutoModelForTokenClassification, AutoTokenizer,  # noqa: F401
                              BitsAndBytesConfig, DataCollatorForTokenClassification, Trainer, TrainingArguments)

    for top in list(STUB_TOP):
        try:
            __import__(top)
            STUB_TOP.discard(top)  # really installed: use it
        except Exception:
            pass
    sys.meta_path.insert(0, _StubFinder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def import_reference(modname: str, max_repairs: int = 60):
    """Import a reference module, shimming names that newer transformers releases no longer export
    (the reference pins transformers==4.31; class bodies that subclass them are never executed by us)."""
    import importlib
    import re
    import transformers.models.esm.modeling_esm as hf_esm
    import transformers.models.llama.modeling_llama as hf_llama

    hf_llama.__dict__.setdefault("add_start_docstrings_to_model_forward", lambda *a, **k: (lambda f: f))
    hf_llama.__dict__.setdefault("LLAMA_INPUTS_DOCSTRING", "")
    for _ in range(max_repairs):
        try:
            return importlib.import_module(modname)
        except NameError as e:
            name = re.search(r"name '(\w+)' is not defined", str(e)).group(1)
            target = hf_esm
        except ImportError as e:
            m = re.search(r"cannot import name '(\w+)' from '([\w.]+)'", str(e))
            if not m:
                raise
            name, target = m.group(1), importlib.import_module(m.group(2))
        if not hasattr(target, name):  # never replace a real class: only fill in what the new release dropped
            setattr(target, name, type(name, (), {"__init__": lambda self, *a, **k: None}))
        if hasattr(target, "__all__") and name not in target.__all__:
            target.__all__.append(name)
        for k in [k for k in sys.modules if k.startswith("procyon.")]:
            if getattr(sys.modules[k], "__spec__", None) is None or k == modname:
                sys.modules.pop(k, None)
        sys.modules.pop(modname, None)
    raise RuntimeError(f"could not import {modname}")
