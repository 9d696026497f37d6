"""TEST INFRASTRUCTURE ONLY - imports the UNMODIFIED reference (FingerRec/OA-Transformer, /root/reference) in the
authoring container so that (a) the oracle restatement in oracle/oracle.py can be validated against it and
(b) golden fixtures can be generated (oracle/make_golden.py -> tests/golden/*.pt).

/root/reference does not exist on the GPU box; nothing in the product, smoke() or bench.py imports this module.

Import-only shims (none touches arithmetic), following SURVEY.md section 8c:
  * timm.models.layers: DropPath -> identity (drop_path_rate = 0 everywhere, video_transformer.py:197),
    to_2tuple, trunc_normal_ = torch.nn.init.trunc_normal_
  * av, decord, humanize, ipdb, matplotlib: empty modules (dataset / util / metric imports)
  * a scratch cwd holding pretrained/distilbert-base-uncased (seeded random weights) and an empty ViT state dict,
    because FrozenInTime.__init__ loads both unconditionally (oa_model.py:27,42).
"""
import os
import sys
import tempfile
import types

REF_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "OATrans"))


def install():
    """Register the import shims and put the reference on sys.path. Idempotent."""
    import transformers  # noqa: F401  must be imported before the timm shim (lazy-module find_spec)
    import torch
    from torch import nn

    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")

        class DropPath(nn.Module):
            def __init__(self, p=0.0):
                super().__init__()
                assert p == 0.0

            def forward(self, x):
                return x

        def to_2tuple(x):
            return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

        layers.DropPath = DropPath
        layers.to_2tuple = to_2tuple
        layers.trunc_normal_ = nn.init.trunc_normal_
        timm.models = models
        models.layers = layers
        import importlib.machinery
        for name, mod in (("timm", timm), ("timm.models", models), ("timm.models.layers", layers)):
            mod.__spec__ = importlib.machinery.ModuleSpec(name, None)
            sys.modules[name] = mod
    for name in ("av", "humanize", "ipdb", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            import importlib.machinery
            m.__spec__ = importlib.machinery.ModuleSpec(name, None)
            sys.modules[name] = m
    if "decord" not in sys.modules:
        import importlib.machinery
        d = types.ModuleType("decord")
        d.__spec__ = importlib.machinery.ModuleSpec("decord", None)
        d.bridge = types.SimpleNamespace(set_bridge=lambda *_a, **_k: None)
        sys.modules["decord"] = d
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    return torch


def scratch_cwd(seed=0):
    """Create (once per process) a scratch directory with the files FrozenInTime.__init__ insists on loading and
    chdir into it. Returns the path."""
    import torch
    from transformers import DistilBertConfig, DistilBertModel

    d = tempfile.mkdtemp(prefix="oat_ref_")
    os.makedirs(os.path.join(d, "pretrained"), exist_ok=True)
    torch.manual_seed(seed)
    DistilBertModel(DistilBertConfig()).save_pretrained(os.path.join(d, "pretrained", "distilbert-base-uncased"))
    torch.save({}, os.path.join(d, "pretrained", "jx_vit_base_p16_224-80ecf9dd.pth"))
    os.chdir(d)
    return d


def load_file_module(name, relpath):
    """Load one reference source file by path (for files without package-relative imports)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
