"""Harness that imports the *reference itself* (/root/reference, read-only, build container
only) so the oracle can be pinned against it and golden vectors generated.

Nothing here is product code and nothing here runs on the GPU box (the reference tree does
not exist there).  Five third-party packages the reference imports are not installed; they are
replaced by minimal import shims that provide only the names the reference touches at import /
construction time.  The shims contain no arithmetic of the hot path.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import torch
import torch.nn as nn

REFERENCE_ROOT = Path("/root/reference")


def reference_available() -> bool:
    return (REFERENCE_ROOT / "model" / "rdeic.py").exists()


def _mod(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


def install_shims() -> None:
    if "pytorch_lightning" in sys.modules and getattr(sys.modules["pytorch_lightning"], "_rdeic_shim", False):
        return
    # ---- pytorch_lightning -------------------------------------------------------------
    pl = _mod("pytorch_lightning")
    pl._rdeic_shim = True

    class LightningModule(nn.Module):
        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def freeze(self):
            for p in self.parameters():
                p.requires_grad = False
            self.eval()

        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

    class Callback:
        pass

    pl.LightningModule = LightningModule
    pl.Callback = Callback
    pl.seed_everything = lambda s, **k: torch.manual_seed(s)
    cb = _mod("pytorch_lightning.callbacks")
    cb.Callback = Callback
    cb.ModelCheckpoint = type("ModelCheckpoint", (Callback,), {})
    pl.callbacks = cb
    ut = _mod("pytorch_lightning.utilities")
    utt = _mod("pytorch_lightning.utilities.types")
    utt.EPOCH_OUTPUT = object
    utt.STEP_OUTPUT = object
    utd = _mod("pytorch_lightning.utilities.distributed")
    utd.rank_zero_only = lambda f: f
    utr = _mod("pytorch_lightning.utilities.rank_zero")
    utr.rank_zero_only = lambda f: f
    ut.types, ut.distributed, ut.rank_zero = utt, utd, utr
    pl.utilities = ut
    # ---- omegaconf -----------------------------------------------------------------------
    oc = _mod("omegaconf")

    class ListConfig(list):
        pass

    class OmegaConf:
        @staticmethod
        def load(path):
            import yaml

            with open(path) as f:
                return yaml.safe_load(f)

        @staticmethod
        def create(obj):
            return obj

    oc.OmegaConf, oc.ListConfig = OmegaConf, ListConfig
    ocl = _mod("omegaconf.listconfig")
    ocl.ListConfig = ListConfig
    oc.listconfig = ocl
    # ---- pyiqa -----------------------------------------------------------------------------
    pq = _mod("pyiqa")
    pq.create_metric = lambda name, **k: (lambda *a, **kk: torch.zeros(()))
    # ---- compressai / torchac: import-time names only (the entropy nets are not exercised) ---
    ca = _mod("compressai")
    cal = _mod("compressai.layers")
    cal.conv3x3 = lambda i, o, stride=1: nn.Conv2d(i, o, kernel_size=3, stride=stride, padding=1)
    cam = _mod("compressai.models")
    cam.CompressionModel = type("CompressionModel", (nn.Module,), {})
    cae = _mod("compressai.entropy_models")
    cae.EntropyModel = type("EntropyModel", (nn.Module,), {})
    cae.GaussianConditional = type("GaussianConditional", (nn.Module,), {"__init__": lambda self, *a, **k: nn.Module.__init__(self)})
    cao = _mod("compressai.ops")
    cao.quantize_ste = lambda x: torch.round(x)
    caa = _mod("compressai.ans")
    caa.BufferedRansEncoder = type("BufferedRansEncoder", (), {})
    caa.RansDecoder = type("RansDecoder", (), {})
    ca.layers, ca.models, ca.entropy_models, ca.ops, ca.ans = cal, cam, cae, cao, caa
    _mod("torchac")


def import_reference():
    """Put /root/reference on sys.path (after the shims) and return its key modules."""
    if not reference_available():
        raise RuntimeError("/root/reference is not present on this machine")
    install_shims()
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))
    import importlib

    return {
        "rdeic": importlib.import_module("model.rdeic"),
        "spaced": importlib.import_module("model.spaced_sampler_relay"),
        "ddim": importlib.import_module("model.ddim_sampler_relay"),
        "ckbd": None,  # utils/ckbd.py needs real compressai arithmetic; not imported here
        "openaimodel": importlib.import_module("ldm.modules.diffusionmodules.openaimodel"),
        "vae": importlib.import_module("ldm.modules.diffusionmodules.model"),
        "autoencoder": importlib.import_module("ldm.models.autoencoder"),
        "attention": importlib.import_module("ldm.modules.attention"),
        "util": importlib.import_module("ldm.modules.diffusionmodules.util"),
    }


def load_config(overrides: dict | None = None) -> dict:
    """configs/model/rdeic.yaml with the offline overrides of SURVEY.md §8c."""
    import copy

    import yaml

    with open(REFERENCE_ROOT / "configs" / "model" / "rdeic.yaml") as f:
        cfg = yaml.safe_load(f)
    p = cfg["params"]
    p["sync_path"] = None
    p["is_refine"] = False
    p["cond_stage_config"] = {"target": "torch.nn.Identity"}
    p["preprocess_config"] = {"target": "torch.nn.Identity"}
    p["unet_config"]["params"]["use_checkpoint"] = False
    p["control_stage_config"]["params"]["use_checkpoint"] = False
    for k, v in (overrides or {}).items():
        d = p
        ks = k.split(".")
        for kk in ks[:-1]:
            d = d[kk]
        d[ks[-1]] = copy.deepcopy(v)
    return cfg


SMALL_OVERRIDES = {
    # reduced-width variant of the same architecture for fast CPU tests
    "unet_config.params.model_channels": 64,
    "unet_config.params.num_head_channels": 16,
    "control_stage_config.params.model_channels": 64,
    "control_stage_config.params.num_head_channels": 16,
    "control_stage_config.params.control_model_ratio": 0.5,
    "control_stage_config.params.hint_channels": 32,
    "control_stage_config.params.context_dim": 64,
    "unet_config.params.context_dim": 64,
    "first_stage_config.params.ddconfig.ch": 32,
}


def reseed_zero_params(model: nn.Module, seed: int = 1234, std: float = 0.02) -> int:
    """zero_module()-initialised tensors make apply_model return exactly 0;
    overwrite every exactly-zero parameter tensor with seeded N(0, std)."""
    g = torch.Generator().manual_seed(seed)
    n = 0
    with torch.no_grad():
        for _, p in model.named_parameters():
            if p.numel() > 0 and torch.count_nonzero(p) == 0:
                p.copy_(torch.randn(p.shape, generator=g) * std)
                n += 1
    return n


def build_reference_model(overrides: dict | None = None, init_seed: int = 231):
    """Instantiate the reference RDEIC (model/rdeic.py:600) on CPU with seeded random weights."""
    mods = import_reference()
    cfg = load_config(overrides)
    from ldm.util import instantiate_from_config

    torch.manual_seed(init_seed)
    model = instantiate_from_config(cfg)
    model.eval()
    reseed_zero_params(model)
    return model, mods


def patch_ddim_register_buffer(ddim_mod) -> None:
    """model/ddim_sampler_relay.py:17-21 forces buffers to 'cuda'; allow CPU for the oracle run."""

    def register_buffer(self, name, attr):
        setattr(self, name, attr)

    ddim_mod.DDIMSampler.register_buffer = register_buffer
