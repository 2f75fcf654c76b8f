"""Stage the unmodified reference for the model-level tests / bench arms, and make it importable.

The reference (nkkrnkl/category-agnostic-pose-estimation) is pure Python with no setup.py / pyproject, so ``pip install
--target baseline/_ref /root/reference`` has nothing to build.  This recipe does what that install would have done:

    python tools/stage_reference.py            # copy /root/reference/{models,util,datasets}/*.py -> baseline/_ref/

``baseline/_ref/`` is git-ignored (never part of the history) but not gpurun-ignored, so it travels to the GPU box,
where ``/root/reference`` does not exist.  Nothing under ``category-agnostic-pose-estimation_b200/`` imports it: it is
used by ``tests/`` (model-level parity through ``patch_reference``), by ``bench.py``'s reference arms, and by nothing
else.

``activate()`` puts the reference on ``sys.path`` (the staged copy, else ``/root/reference`` when that exists) with the
three shims the survey recorded for a box without network:

* ``pycocotools`` — imported by ``datasets/mp100_cape.py:9`` at package-import time; no COCO file is ever opened here;
* ``timm.layers`` (``DropPath``, ``Mlp``) — imported by ``models/bixattn.py:3-4``, a module CAPE does not execute;
* ``torchvision.models.resnet50(weights=None)`` — ``models/backbone.py:75-78`` asks for the ImageNet weights, which
  would be a download; the benchmark uses random-init weights of the same architecture.
"""
from __future__ import annotations

import os
import shutil
import sys
import types

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCE = "/root/reference"
STAGED = os.path.join(REPO, "baseline", "_ref")
PACKAGES = ("models", "util", "datasets")
_activated = None


def stage(force: bool = False) -> str | None:
    """Copy the reference's Python packages into baseline/_ref.  Returns the staged path, or None when the reference is
    not available on this machine (e.g. the GPU box — the staged copy made in the build container is used there)."""
    if not os.path.isdir(SOURCE):
        return STAGED if os.path.isdir(os.path.join(STAGED, "models")) else None
    os.makedirs(STAGED, exist_ok=True)
    for pkg in PACKAGES:
        src, dst = os.path.join(SOURCE, pkg), os.path.join(STAGED, pkg)
        if os.path.isdir(dst):
            if not force and _same_tree(src, dst):
                continue
            shutil.rmtree(dst)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for name in ("category_splits.json",):
        if os.path.exists(os.path.join(SOURCE, name)):
            shutil.copy2(os.path.join(SOURCE, name), os.path.join(STAGED, name))
    with open(os.path.join(STAGED, "STAGED_FROM"), "w") as f:
        f.write(SOURCE + "\n")
    return STAGED


def _same_tree(a: str, b: str) -> bool:
    for root, _, files in os.walk(a):
        if "__pycache__" in root:
            continue
        for fn in files:
            if fn.endswith(".pyc"):
                continue
            pa = os.path.join(root, fn)
            pb = os.path.join(b, os.path.relpath(pa, a))
            if not os.path.exists(pb) or os.path.getsize(pa) != os.path.getsize(pb) \
                    or os.path.getmtime(pa) > os.path.getmtime(pb):
                return False
    return True


def root() -> str | None:
    """Directory the reference would be imported from: the staged copy if present, else /root/reference."""
    if os.path.isdir(os.path.join(STAGED, "models")):
        return STAGED
    if os.path.isdir(os.path.join(SOURCE, "models")):
        return SOURCE
    return None


def available() -> bool:
    return root() is not None


def _install_shims():
    import torch
    if "pycocotools" not in sys.modules:
        pc, pcc, pcm = types.ModuleType("pycocotools"), types.ModuleType("pycocotools.coco"), types.ModuleType("pycocotools.mask")

        class COCO:   # never instantiated: the benchmark builds synthetic episodes, not an MP-100 dataset
            def __init__(self, *a, **k):
                raise RuntimeError("pycocotools is not installed (shim from tools/stage_reference.py)")
        pcc.COCO = COCO
        pc.coco, pc.mask = pcc, pcm
        sys.modules.update({"pycocotools": pc, "pycocotools.coco": pcc, "pycocotools.mask": pcm})
    try:
        import timm.layers  # noqa: F401
    except Exception:
        tm, tml = types.ModuleType("timm"), types.ModuleType("timm.layers")

        class DropPath(torch.nn.Identity):
            def __init__(self, *a, **k):
                super().__init__()

        class Mlp(torch.nn.Module):
            def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=torch.nn.GELU, drop=0.0, **k):
                super().__init__()
                self.fc1 = torch.nn.Linear(in_features, hidden_features or in_features)
                self.act = act_layer()
                self.fc2 = torch.nn.Linear(hidden_features or in_features, out_features or in_features)

            def forward(self, x):
                return self.fc2(self.act(self.fc1(x)))
        tml.DropPath, tml.Mlp, tm.layers = DropPath, Mlp, tml
        sys.modules.update({"timm": tm, "timm.layers": tml})
    import torchvision
    if not getattr(torchvision.models.resnet50, "_cape_no_download", False):
        orig = torchvision.models.resnet50

        def resnet50(*args, **kwargs):
            kwargs.pop("pretrained", None)
            kwargs["weights"] = None            # no network: random init of the same architecture
            return orig(*args, **kwargs)
        resnet50._cape_no_download = True
        torchvision.models.resnet50 = resnet50


def activate() -> str:
    """Make ``import models`` / ``util`` / ``datasets`` resolve to the reference.  Returns the directory used."""
    global _activated
    if _activated:
        return _activated
    r = root()
    if r is None:
        raise RuntimeError("reference not available: run `python tools/stage_reference.py` where /root/reference exists")
    _install_shims()
    if r not in sys.path:
        sys.path.insert(0, r)
    _activated = r
    return r


def build_cape_model(device="cpu", extra_args=(), seed=0):
    """The model of BASELINE.json's configs: parser defaults + --use_geometric_encoder --use_gcn_preenc
    (train_mp100_cape_cola.ipynb cell 23), DiscreteTokenizerV2(44 bins, 200 tokens), ResNet-50 random init.
    Returns (model, criterion, args, tokenizer) — all objects of the UNMODIFIED reference."""
    import torch
    activate()
    from models.train_cape_episodic import get_args_parser
    from models import build_model
    from models.cape_model import build_cape_model as _build_cape
    from models.cape_losses import build_cape_criterion
    from datasets.discrete_tokenizer import DiscreteTokenizerV2
    args = get_args_parser().parse_args(["--use_geometric_encoder", "--use_gcn_preenc", "--device", str(device),
                                         *extra_args])
    tok = DiscreteTokenizerV2(num_bins=44, seq_len=args.seq_len, add_cls=False)   # int(sqrt(vocab_size 2000)), mp100_cape.py:116-121
    torch.manual_seed(seed)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):        # the builders print banners; keep bench stdout to one line
        built = build_model(args, tokenizer=tok)
        base = built[0] if isinstance(built, tuple) else built
        model = _build_cape(args, base)
        criterion = build_cape_criterion(args, num_classes=3)
    return model.to(device), criterion.to(device), args, tok


def build_optimizer(model, args):
    """AdamW with the reference's two parameter groups (models/train_cape_episodic.py:527-538)."""
    import torch
    groups = [{"params": [p for n, p in model.named_parameters() if "backbone" not in n and p.requires_grad]},
              {"params": [p for n, p in model.named_parameters() if "backbone" in n and p.requires_grad],
               "lr": args.lr_backbone}]
    return torch.optim.AdamW(groups, lr=args.lr, weight_decay=args.weight_decay)


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
