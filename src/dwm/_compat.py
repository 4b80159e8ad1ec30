"""Marker base classes for `isinstance(model, diffusers.SD3Transformer2DModel)` checks.

The reference pipeline distinguishes SD-3 from SD-2.1 models with isinstance tests
against diffusers classes (src/dwm/pipelines/ctsd.py:186,205,450,888,...).  diffusers is
not a dependency of this implementation; when it is importable its real classes are
used as the markers so mixed code keeps working, otherwise minimal stand-ins are
registered under the `diffusers` module name.
"""
import sys
import types

import torch

try:  # pragma: no cover - diffusers is absent in the build / GPU images
    import diffusers as _diffusers
    SD3Transformer2DModelMarker = _diffusers.SD3Transformer2DModel
    UNetSpatioTemporalConditionModelMarker = \
        _diffusers.UNetSpatioTemporalConditionModel
    HAVE_DIFFUSERS = True
except Exception:  # noqa: BLE001
    HAVE_DIFFUSERS = False

    class SD3Transformer2DModelMarker(torch.nn.Module):
        pass

    class UNetSpatioTemporalConditionModelMarker(torch.nn.Module):
        pass

    _stub = types.ModuleType("diffusers")
    _stub.SD3Transformer2DModel = SD3Transformer2DModelMarker
    _stub.UNetSpatioTemporalConditionModel = \
        UNetSpatioTemporalConditionModelMarker
    _stub.__dwm_b200_stub__ = True
    sys.modules.setdefault("diffusers", _stub)
