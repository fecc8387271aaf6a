"""Capability probes feeding ``CompilationConfig`` defaults
(reference: /root/reference/src/sfast/utils/gpu_device.py:4-15)."""
import torch


def device_has_tensor_core():
    if torch.cuda.is_available():
        major, _ = torch.cuda.get_device_capability()
        return major >= 7
    return False


def device_has_capability(major, minor):
    if torch.cuda.is_available():
        return tuple(torch.cuda.get_device_capability()) >= (major, minor)
    return False
