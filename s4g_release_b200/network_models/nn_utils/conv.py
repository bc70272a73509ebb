"""1x1 convolution blocks with the reference's parameter layout (``conv``/``bn`` sub-modules, so the
state_dict keys ``*.conv.weight`` / ``*.bn.{weight,bias,running_mean,running_var,num_batches_tracked}``
match inference/grasp_proposal/network_models/nn_utils/conv.py:6-86 and its checkpoints load with
strict=True).  conv has no bias when batch-norm follows (reference :24, :64)."""
from torch import nn


class _ConvBNReLU(nn.Module):
    _conv_cls = None
    _bn_cls = None

    def __init__(self, in_channels, out_channels, kernel_size, relu=True, bn=True, bn_momentum=0.1, **kwargs):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.conv = self._conv_cls(in_channels, out_channels, kernel_size, bias=(not bn), **kwargs)
        self.bn = self._bn_cls(out_channels, momentum=bn_momentum) if bn else None
        self.relu = nn.ReLU(inplace=True) if relu else None
        self.init_weights()

    def forward(self, x):
        x = self.conv(x)
        if self.bn is not None:
            x = self.bn(x)
        return x if self.relu is None else self.relu(x)

    def init_weights(self, init_fn=None):
        if init_fn is not None:
            init_fn(self.conv)
        if self.bn is not None:  # reference nn_utils/init.py:4-8
            nn.init.ones_(self.bn.weight)
            nn.init.zeros_(self.bn.bias)


class Conv1d(_ConvBNReLU):
    _conv_cls = nn.Conv1d
    _bn_cls = nn.BatchNorm1d


class Conv2d(_ConvBNReLU):
    _conv_cls = nn.Conv2d
    _bn_cls = nn.BatchNorm2d
