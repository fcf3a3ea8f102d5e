"""vgtk.spconv.modules -- legacy S^2-anchor ZPConv modules (reference: vgtk/vgtk/spconv/modules.py).

The shipped models only use the SO(3) modules in vgtk.so3conv; the legacy ZPConv family
(BasicZPConv / IntraZPConv / InterZPConv / AnchorProp) is outside the hot-path scope
 and raises on construction."""
import torch.nn as nn


class _OutOfScope(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError(f"{type(self).__name__}: legacy S^2 ZPConv is outside the B200 hot-path scope "
                                  "(no shipped model builds it)")


class BasicZPConv(_OutOfScope):
    pass


class IntraZPConv(_OutOfScope):
    pass


class InterZPConv(_OutOfScope):
    pass


class AnchorProp(_OutOfScope):
    pass
