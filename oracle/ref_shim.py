"""Import the UNMODIFIED reference DVC modules from /root/reference.  TEST INFRASTRUCTURE ONLY.

Only usable where the reference checkout exists (the build container); never on the GPU box.
Shims are import-time / device plumbing only and touch no arithmetic:
  1. stub modules for ``torchac`` (net.py:15; used only when calrealbits=True) and
     ``matplotlib`` (flowlib.py:10-11, imported by basics.py:18);
  2. cwd is switched to the reference root while the model is built so that
     ``loadweightformnp`` (endecoder.py:122-139) finds ``DVC/flow_pretrain_np``;
  3. on CPU, ``endecoder.torch_warp`` is replaced by a device-agnostic copy of the same
     four statements (``device.index`` is None on CPU, endecoder.py:54-57).
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("FVC_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "DVC", "net.py"))


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def _stub(name):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__dict__["__stub__"] = True
        sys.modules[name] = m
    return sys.modules[name]


_grids = {}


def _torch_warp_any_device(tensorInput, tensorFlow):
    # same statements as endecoder.py:52-67, with the per-device grid cache keyed on the device
    # object instead of device.index
    key = (str(tensorInput.device), str(tensorFlow.size()))
    if key not in _grids:
        hor = torch.linspace(-1.0, 1.0, tensorFlow.size(3)).view(1, 1, 1, tensorFlow.size(3)).expand(
            tensorFlow.size(0), -1, tensorFlow.size(2), -1)
        ver = torch.linspace(-1.0, 1.0, tensorFlow.size(2)).view(1, 1, tensorFlow.size(2), 1).expand(
            tensorFlow.size(0), -1, -1, tensorFlow.size(3))
        _grids[key] = torch.cat([hor, ver], 1).to(tensorInput.device)
    tensorFlow = torch.cat([tensorFlow[:, 0:1, :, :] / ((tensorInput.size(3) - 1.0) / 2.0),
                            tensorFlow[:, 1:2, :, :] / ((tensorInput.size(2) - 1.0) / 2.0)], 1)
    return torch.nn.functional.grid_sample(input=tensorInput, grid=(_grids[key] + tensorFlow).permute(0, 2, 3, 1),
                                           mode='bilinear', padding_mode='border')


def load_reference():
    """Returns the reference ``DVC.net`` module (imported once)."""
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REF_ROOT)
    _stub("torchac")
    mpl = _stub("matplotlib")
    mpl.colors = _stub("matplotlib.colors")
    mpl.pyplot = _stub("matplotlib.pyplot")
    mpl.colors.hsv_to_rgb = lambda *a, **k: None
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with _cwd(REF_ROOT):
        import DVC.net as refnet  # noqa
        import DVC.subnet.endecoder as endec
    endec.torch_warp = _torch_warp_any_device
    return refnet


def build_reference_model(state_dict=None):
    """Reference VideoCompressor().eval() on CPU; optionally loaded with ``state_dict``."""
    refnet = load_reference()
    with _cwd(REF_ROOT):
        model = refnet.VideoCompressor()
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    return model.eval()


def run_reference_with_capture(model, cur, ref):
    """Runs reference forward and captures the intermediates via forward hooks (no code changes)."""
    cap = {}
    hooks = []

    def save(name):
        def fn(mod, inp, out):
            cap[name] = out.detach().clone()
        return fn

    def save_in(name):
        def fn(mod, inp):
            cap[name] = inp[0].detach().clone()
        return fn

    hooks.append(model.opticFlow.register_forward_hook(save("estmv")))
    hooks.append(model.mvEncoder.register_forward_hook(save("mvfeature")))
    hooks.append(model.mvDecoder.register_forward_pre_hook(save_in("quant_mv")))
    hooks.append(model.mvDecoder.register_forward_hook(save("mv_hat")))
    hooks.append(model.warpnet.register_forward_pre_hook(save_in("warp_in")))
    hooks.append(model.warpnet.register_forward_hook(save("warpnet_out")))
    hooks.append(model.resEncoder.register_forward_hook(save("feature")))
    hooks.append(model.respriorEncoder.register_forward_hook(save("z")))
    hooks.append(model.respriorDecoder.register_forward_pre_hook(save_in("z_hat")))
    hooks.append(model.respriorDecoder.register_forward_hook(save("sigma")))
    hooks.append(model.resDecoder.register_forward_pre_hook(save_in("feat_hat")))
    hooks.append(model.resDecoder.register_forward_hook(save("recon_res")))
    with torch.no_grad():
        out = model(cur, ref)
    for h in hooks:
        h.remove()
    cap["warpframe"] = cap["warp_in"][:, 0:3].clone()
    cap["prediction"] = cap["warpnet_out"] + cap["warpframe"]
    cap["recon"] = cap["prediction"] + cap["recon_res"]
    del cap["warp_in"], cap["warpnet_out"]
    return out, cap


class _AnyModule(types.ModuleType):
    """Stub package for the reference's un-vendored imports (compressai, pytorch_msssim): every attribute is a
    placeholder nn.Module subclass, which is enough for ``models.py`` to define (not run) the codecs built on them."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (torch.nn.Module,), {"__init__": lambda self, *a, **k: torch.nn.Module.__init__(self)})
        setattr(self, name, cls)
        return cls


def load_reference_models():
    """Imports the reference ``models.py`` (GOP driver, LSVC, ...) with import-time stubs only: compressai and
    pytorch_msssim placeholders (none of them is touched by the DVC / LSVC-128 paths), torchac, and the
    device-agnostic copy of ``torch_warp`` (models.py:732-741 has the same 4 statements as endecoder.py:52-67)."""
    load_reference()
    names = ["compressai", "compressai.entropy_models", "compressai.models", "compressai.models.waseda",
             "compressai.models.video", "compressai.models.video.google", "compressai.models.utils",
             "compressai.layers", "compressai.ops", "compressai.zoo", "compressai.ans", "pytorch_msssim"]
    for name in names:
        if name not in sys.modules:
            sys.modules[name] = _AnyModule(name)
    for name in names:
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, sys.modules[name])
    sys.modules["torchac"] = _AnyModule("torchac")
    with _cwd(REF_ROOT):
        import models as refmodels  # noqa
    refmodels.torch_warp = _torch_warp_any_device
    refmodels.flow_warp = _torch_warp_any_device
    return refmodels
