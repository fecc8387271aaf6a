"""Capability probes feeding ``CompilationConfig`` defaults (same public names as the reference's
``sfast.utils.gpu_device``).  This build only ever runs its fast path on sm_100, but the config
defaults must still evaluate on any box, including one without a GPU."""
import torch

_NO_GPU = (0, 0)


def _capability():
    """(major, minor) of the current CUDA device, (0, 0) without one."""
    return tuple(torch.cuda.get_device_capability()) if torch.cuda.is_available() else _NO_GPU


def device_has_capability(major, minor):
    return _capability() >= (major, minor)


def device_has_tensor_core():
    return device_has_capability(7, 0)
