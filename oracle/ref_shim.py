"""TEST INFRASTRUCTURE ONLY -- makes the reference's own Python classes importable on CPU.

Only usable where ``/root/reference`` is mounted (the build container).  It is used by
``tests/golden/make_golden.py`` to generate the committed fixtures and by the optional
``tests/test_oracle_vs_reference.py`` cross-check; nothing that runs on the GPU box imports it.

Recipe: stub the absent third-party packages in ``sys.modules``, alias the
``modulus.*`` modules to the reference's local copies under ``src/models/sfno/`` and provide
``torch_harmonics`` through the restatement in ``oracle/harmonics.py``.  The reference sources are
imported where they lie; nothing is copied.
"""
from __future__ import annotations

import contextlib
import importlib.machinery
import importlib.util
import inspect
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("SFNO_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "models", "sfno", "sfnonet.py"))


def _stub(name: str, is_pkg: bool = False, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__spec__ = importlib.machinery.ModuleSpec(name, loader=None, is_package=is_pkg)
    if is_pkg:
        mod.__path__ = []
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, mod)
    return mod


class _AttrDict(dict):
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


class _LightningModule(nn.Module):
    """Just enough of ``pytorch_lightning.LightningModule`` for ``BaseModel`` (``_base_model.py:23-112``)."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        self._hparams = _AttrDict()

    @property
    def hparams(self):
        return self._hparams

    def save_hyperparameters(self, *args, ignore=None, **kwargs):
        ignore = set(ignore or [])
        frame = inspect.currentframe().f_back
        local_vars = frame.f_locals
        for k, v in local_vars.items():
            if k in ("self", "__class__") or k in ignore or k.startswith("_"):
                continue
            if k == "kwargs" and isinstance(v, dict):
                for kk, vv in v.items():
                    if kk not in ignore:
                        self._hparams[kk] = vv
                continue
            self._hparams[k] = v

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def log(self, *a, **k):
        pass

    def log_dict(self, *a, **k):
        pass


def _alias_file(mod_name: str, path: str) -> types.ModuleType:
    spec = importlib.util.spec_from_file_location(mod_name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[mod_name] = mod
    parent, _, child = mod_name.rpartition(".")
    if parent in sys.modules:
        setattr(sys.modules[parent], child, mod)
    spec.loader.exec_module(mod)
    return mod


_INSTALLED = False


def install() -> None:
    """Idempotently install the stubs and put the reference on ``sys.path``."""
    global _INSTALLED
    if _INSTALLED:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    from oracle import harmonics

    def have(name):
        try:
            return importlib.util.find_spec(name) is not None
        except (ImportError, ValueError):
            return False

    # --- lightning ---------------------------------------------------------------------------
    if not have("pytorch_lightning"):
        pl = _stub("pytorch_lightning", True, LightningModule=_LightningModule,
                   LightningDataModule=object, Trainer=object, Callback=object,
                   seed_everything=lambda s, **k: torch.manual_seed(s))
        _stub("pytorch_lightning.utilities", True, rank_zero_only=lambda f: f)
        _stub("pytorch_lightning.utilities.types", False, STEP_OUTPUT=object, EVAL_DATALOADERS=object,
              TRAIN_DATALOADERS=object)
        _stub("pytorch_lightning.callbacks", True, Callback=object, ModelCheckpoint=object)
        _stub("pytorch_lightning.loggers", True, WandbLogger=object, Logger=object)
        _stub("pytorch_lightning.loggers.wandb", False, WandbLogger=object)
        pl.utilities = sys.modules["pytorch_lightning.utilities"]

    # --- hydra / omegaconf -------------------------------------------------------------------
    if not have("hydra"):
        def _instantiate(cfg, *a, **k):
            raise RuntimeError("hydra.utils.instantiate is not available in the oracle shim")
        _stub("hydra", True)
        _stub("hydra.utils", False, instantiate=_instantiate, get_original_cwd=os.getcwd)
        _stub("hydra.core", True)
        _stub("hydra.core.hydra_config", False, HydraConfig=object)
    if not have("omegaconf"):
        class _OmegaConf:
            @staticmethod
            def to_container(x, **k):
                return x

            @staticmethod
            def create(x=None, **k):
                return x if x is not None else {}

            @staticmethod
            def is_config(x):
                return False

            @staticmethod
            def register_new_resolver(*a, **k):
                pass

            @staticmethod
            def set_struct(*a, **k):
                pass

        _stub("omegaconf", True, DictConfig=dict, ListConfig=list, OmegaConf=_OmegaConf, open_dict=contextlib.nullcontext)
        _stub("omegaconf.errors", False, ConfigAttributeError=AttributeError)

    # --- data containers ---------------------------------------------------------------------
    if not have("xarray"):
        _stub("xarray", True, DataArray=type("DataArray", (), {}), Dataset=type("Dataset", (), {}))
    if not have("tensordict"):
        _stub("tensordict", True, TensorDict=dict, TensorDictBase=dict)
    if not have("dacite"):
        _stub("dacite", True)
    if not have("netCDF4"):
        _stub("netCDF4", True)
    if not have("matplotlib"):
        _stub("matplotlib", True)
        _stub("matplotlib.pyplot", False)

    # --- tensor factorisation packages (only touched at import time) ---------------
    if not have("tensorly"):
        _stub("tensorly", True, set_backend=lambda *a, **k: None, ndim=lambda t: t.ndim, einsum=torch.einsum)
    if not have("tltorch"):
        _stub("tltorch", True)
        _stub("tltorch.factorized_tensors", True)
        _stub("tltorch.factorized_tensors.core", False, FactorizedTensor=type("FactorizedTensor", (), {}))

    # --- torch_harmonics = restatement ----------------------------------------------------------
    if not have("torch_harmonics"):
        th = _stub("torch_harmonics", True, RealSHT=harmonics.RealSHT, InverseRealSHT=harmonics.InverseRealSHT)
        th.__all__ = ["RealSHT", "InverseRealSHT"]
        _stub("torch_harmonics.distributed", False,
              DistributedRealSHT=type("DistributedRealSHT", (), {}),
              DistributedInverseRealSHT=type("DistributedInverseRealSHT", (), {}),
              init=lambda *a, **k: None)

    # --- modulus -> the reference's own local copies ----------------------------------------------
    if not have("modulus"):
        sfno_dir = os.path.join(REFERENCE_ROOT, "src", "models", "sfno")
        for pkg in ("modulus", "modulus.models", "modulus.models.sfno", "modulus.utils", "modulus.utils.sfno",
                    "modulus.utils.sfno.distributed"):
            _stub(pkg, True)
        _stub("modulus.utils.sfno.logging_utils", False, disable_logging=contextlib.nullcontext)
        for name in ("initialization", "activations", "contractions", "factorizations"):
            _alias_file(f"modulus.models.sfno.{name}", os.path.join(sfno_dir, f"{name}.py"))
        for name in ("comm", "helpers", "mappings"):
            _alias_file(f"modulus.utils.sfno.distributed.{name}", os.path.join(sfno_dir, "distributed", f"{name}.py"))
    _INSTALLED = True


def reference_sfno_class():
    install()
    from src.models.sfno.sfnonet import SphericalFourierNeuralOperatorNet

    return SphericalFourierNeuralOperatorNet


def build_reference_sfno(*, num_input_channels, num_output_channels, num_conditional_channels, spatial_shape,
                         seed=0, min_max_time=(0, 5), **model_kwargs):
    """Instantiate the reference SFNO with its own initialisers under ``torch.manual_seed(seed)``."""
    cls = reference_sfno_class()
    torch.manual_seed(seed)
    model = cls(
        num_input_channels=num_input_channels,
        num_output_channels=num_output_channels,
        num_output_channels_raw=num_output_channels,
        num_conditional_channels=num_conditional_channels,
        spatial_shape_in=tuple(spatial_shape),
        spatial_shape_out=tuple(spatial_shape),
        loss_function=None,
        verbose=False,
        **model_kwargs,
    )
    if model_kwargs.get("with_time_emb", False):
        model.set_min_max_time(*min_max_time)
    return model.eval()
