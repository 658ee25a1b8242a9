"""Generate ``tests/golden/rollout_glue.pt`` from the reference's OWN step-glue classes.

Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden_rollout.py

Loaded unchanged, by path: ``src/ace_inference/core/prescriber.py`` (imports torch only), ``normalizer.py`` (its module
imports ``netCDF4`` and ``src.ace_inference.core.device`` are stubbed -- neither is touched by ``StandardNormalizer``) and
``packer.py`` (``tensordict`` stubbed).  For a few seeded cases the fixture stores a dict of named raw fields, the
per-variable means / standard deviations and what the reference's ``run_on_batch_multistep`` does with them around one
call of the sampler (``stepper_multistep.py:365-427``):

    normalize -> pack(in_names)                                   = ``packed_norm``   (the sampler's input)
    unpack(gen) -> Prescriber(data_t, gen_norm, target_norm)      = ``gen_prescribed`` (seeds the next step)
    denormalize                                                    = ``gen_denorm``     (what is recorded)

``gen`` stands in for the sampler's normalised prediction (seeded noise): the glue never looks inside it.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import types

import torch

OUT = os.path.dirname(os.path.abspath(__file__))
CORE = "/root/reference/src/ace_inference/core"

CASES = {
    # prescribed variable, mask variable, mask value, interpolate
    "mask_int": dict(names=["PRESsfc", "T_7", "surface_temperature", "ocean_fraction"], out=["PRESsfc", "T_7", "surface_temperature"],
                     prescribed="surface_temperature", mask="ocean_fraction", mask_value=1, interpolate=False, B=3, H=6, W=12, seed=0),
    "interp": dict(names=["a", "b", "sst", "frac"], out=["a", "sst", "b"], prescribed="sst", mask="frac", mask_value=1,
                   interpolate=True, B=2, H=5, W=8, seed=1),
    "mask_zero": dict(names=["u", "v", "w", "m"], out=["w", "u", "v"], prescribed="w", mask="m", mask_value=0, interpolate=False,
                      B=1, H=4, W=7, seed=2),
}


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_glue():
    _stub("netCDF4")
    _stub("tensordict", TensorDict=dict)
    for pkg in ("src", "src.ace_inference", "src.ace_inference.core"):
        if pkg not in sys.modules:
            m = _stub(pkg)
            m.__path__ = []
    _stub("src.ace_inference.core.device", get_device=lambda: torch.device("cpu"))
    return (_load("ref_prescriber", os.path.join(CORE, "prescriber.py")), _load("ref_normalizer", os.path.join(CORE, "normalizer.py")),
            _load("ref_packer", os.path.join(CORE, "packer.py")))


def main():
    presc_mod, norm_mod, pack_mod = load_reference_glue()
    out = {}
    for name, c in CASES.items():
        g = torch.Generator().manual_seed(c["seed"])
        B, H, W = c["B"], c["H"], c["W"]
        data, target = {}, {}
        for i, n in enumerate(c["names"]):
            if n == c["mask"]:
                frac = torch.rand(B, H, W, generator=g)
                data[n] = frac if c["interpolate"] else torch.where(frac > 0.55, torch.ones(()), torch.where(frac < 0.3, torch.zeros(()), frac))
            else:
                data[n] = torch.randn(B, H, W, generator=g) * (1.0 + i) * 50.0 + 250.0 * i
            target[n] = data[n] + torch.randn(B, H, W, generator=g) * 3.0
        target[c["mask"]] = data[c["mask"]]
        means = {n: torch.tensor(float(250.0 * i + 0.5)) for i, n in enumerate(c["names"])}
        stds = {n: torch.tensor(float((1.0 + i) * 40.0)) for i, n in enumerate(c["names"])}
        normalizer = norm_mod.StandardNormalizer(means=means, stds=stds)
        in_packer, out_packer = pack_mod.Packer(c["names"]), pack_mod.Packer(c["out"])
        prescriber = presc_mod.Prescriber(c["prescribed"], c["mask"], c["mask_value"], interpolate=c["interpolate"])
        # way in: normalize + pack
        packed_norm = in_packer.pack(normalizer.normalize(data), axis=-3)
        # the sampler's normalised prediction (stand-in) and the way out
        gen = torch.randn(B, len(c["out"]), H, W, generator=g)
        target_norm = normalizer.normalize(target)
        # (the TensorDict stand-in is a plain dict: drop its ``batch_size`` entry)
        gen_norm = {k: v for k, v in out_packer.unpack(gen.clone(), axis=-3).items() if torch.is_tensor(v)}
        gen_norm = prescriber(target, gen_norm, {k: target_norm[k] for k in c["out"]})
        gen_prescribed = out_packer.pack(gen_norm, axis=-3)
        gen_denorm = out_packer.pack(normalizer.denormalize(gen_norm), axis=-3)
        out[name] = dict(spec=c, data=data, target=target, means={k: float(v) for k, v in means.items()},
                         stds={k: float(v) for k, v in stds.items()}, gen=gen, packed_norm=packed_norm,
                         gen_prescribed=gen_prescribed, gen_denorm=gen_denorm)
        print(name, tuple(packed_norm.shape), tuple(gen_prescribed.shape),
              "prescribed points:", int((gen_prescribed != gen).sum()))
    torch.save(out, os.path.join(OUT, "rollout_glue.pt"))


if __name__ == "__main__":
    main()
