"""network_models/functions/functions.py of the reference — the pieces the PointNet2 sibling model needs."""
import torch


def toRotMatrix(repre6d):
    """6-D rotation representation (B, 6, N) -> flattened rotation matrices (B, 9, N)  (functions.py:179-190):
    Gram-Schmidt of the two 3-vectors, third axis by cross product; element (b, 3 i + j, n) = R[b, n][i, j]."""
    a1, a2 = repre6d[:, :3, :], repre6d[:, 3:6, :]
    b1 = a1 / torch.norm(a1, dim=1, keepdim=True)
    b2 = a2 - (a2 * b1).sum(dim=1, keepdim=True) * b1
    b2 = b2 / torch.norm(b2, dim=1, keepdim=True)
    b3 = torch.cross(b1, b2, dim=1)
    return torch.stack([b1, b2, b3], dim=2).contiguous().view(repre6d.shape[0], 9, -1)
