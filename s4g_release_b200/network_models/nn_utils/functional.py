"""Loss helpers of the reference's nn_utils/functional.py that the PN2 family's losses call."""
import torch
import torch.nn.functional as F


def smooth_cross_entropy(input, target, label_smoothing, weight=None):
    """Label-smoothed cross entropy as the reference defines it (nn_utils/functional.py:91-114): the target
    distribution is (1 - eps) on the label plus eps / C everywhere, class weights multiply the per-class terms and
    the mean is over ROWS — unlike torch's ``F.cross_entropy(weight=..., label_smoothing=...)``, which normalises by
    the summed weight of the labels.  input (N, C) logits, target (N,) int64."""
    if input.dim() != 2 or target.dim() != 1:
        raise ValueError("smooth_cross_entropy expects (N, C) logits and (N,) labels")
    eps = float(label_smoothing)
    c = input.shape[1]
    soft = torch.full_like(input, eps / c)
    soft.scatter_add_(1, target.unsqueeze(1), torch.full_like(input[:, :1], 1.0 - eps))
    terms = soft * F.log_softmax(input, dim=1)
    if weight is not None:
        terms = terms * weight.unsqueeze(0)
    return -terms.sum(dim=1).mean()
