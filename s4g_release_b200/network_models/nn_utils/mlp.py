"""SharedMLP — a stack of 1x1 conv + BN + ReLU blocks shared over 1 or 2 trailing dimensions.
Constructor arguments, ModuleList indexing (state_dict keys ``<prefix>.<j>.conv.weight`` …) and the
training-time dropout follow inference/grasp_proposal/network_models/nn_utils/mlp.py:55-118."""
import torch.nn.functional as F
from torch import nn

from .conv import Conv1d, Conv2d


class SharedMLP(nn.ModuleList):
    def __init__(self, in_channels, mlp_channels, ndim=1, dropout_prob=0.0, bn=True, bn_momentum=0.1):
        super().__init__()
        if ndim not in (1, 2):
            raise ValueError('SharedMLP only supports ndim=(1, 2).')
        assert dropout_prob >= 0.0
        self.in_channels = in_channels
        self.out_channels = mlp_channels[-1]
        self.ndim = ndim
        self.dropout_prob = dropout_prob
        block = Conv1d if ndim == 1 else Conv2d
        c = in_channels
        for out_channels in mlp_channels:
            self.append(block(c, out_channels, 1, relu=True, bn=bn, bn_momentum=bn_momentum))
            c = out_channels

    def forward(self, x):
        drop = F.dropout if self.ndim == 1 else F.dropout2d
        for block in self:
            x = block(x)
            if self.training and self.dropout_prob > 0.0:
                x = drop(x, p=self.dropout_prob, training=True)
        return x

    def init_weights(self, init_fn=None):
        for block in self:
            block.init_weights(init_fn)

    def extra_repr(self):
        return 'dropout_prob={}'.format(self.dropout_prob) if self.dropout_prob > 0.0 else ''
