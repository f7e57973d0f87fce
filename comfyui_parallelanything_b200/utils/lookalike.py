"""Test helper: rebuild a module tree out of FOREIGN classes (as if it came from ComfyUI's own ``comfy.ldm.*`` code):
same attribute names, same parameters, same leaf torch layers - but none of this repository's model classes, no
``params`` dataclass, no ``pa_family`` marker.  What structural recognition (exec/recognize.py) must cope with."""
import torch
import torch.nn as nn

KEEP_ATTRS = ("heads", "dim_head", "num_heads")      # what ComfyUI's attention layers carry as plain ints


def launder(m: nn.Module, keep=KEEP_ATTRS, forward_from=None) -> nn.Module:
    if type(m).__module__.startswith("torch.nn"):
        if isinstance(m, nn.Sequential):
            return nn.Sequential(*[launder(c, keep) for c in m])
        if isinstance(m, nn.ModuleList):
            return nn.ModuleList([launder(c, keep) for c in m])
        return m
    if isinstance(m, nn.Sequential):                    # e.g. TimestepEmbedSequential
        return nn.Sequential(*[launder(c, keep) for c in m])
    if isinstance(m, nn.GroupNorm):                     # GroupNorm32-style subclasses
        g = nn.GroupNorm(m.num_groups, m.num_channels, eps=m.eps)
        g.weight, g.bias = m.weight, m.bias
        return g
    cls = type("Comfy" + type(m).__name__, (nn.Module,), {})
    new = cls()
    for name, p in m._parameters.items():
        new.register_parameter(name, p)
    for name, b in m._buffers.items():
        new.register_buffer(name, b)
    for name, child in m.named_children():
        setattr(new, name, launder(child, keep))
    for k in keep:
        if k in m.__dict__:
            setattr(new, k, m.__dict__[k])
    if forward_from is not None:
        new.forward = forward_from
    return new
