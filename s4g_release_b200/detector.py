"""GraspDetector on the device: the reference's ``GraspDetector.eval`` / ``detect`` flow (grasp_detector.py:107-262)
assembled from this package's pieces — pre-processing (preprocess.py), the fused PN2_CLS forward
(network_models + engine.py) and the device-side post-processing (postprocess.py).  Same call signature and return
values as the reference's ``detect`` (numpy ``(k,4,4)`` float64 poses in the camera frame and ``(k,)`` float64
scores); config / checkpoint / logging / visualisation of the reference class are out of scope.
"""
import numpy as np
import torch

from .postprocess import CAMERA2BASE, GraspPostProcessor
from .preprocess import NUM_INPUT, GraspPreProcessor


class GraspDetector:
    def __init__(self, model=None, state_dict=None, camera2base=CAMERA2BASE, num_input=NUM_INPUT, device="cuda"):
        """model: a ``PointNet2`` (default: PN2_CLS, the ``curvature_model`` configuration); state_dict: optional
        parameters in the reference's key layout (the checkpoint's ``"model"`` entry, utils/checkpoint.py:37-44)."""
        from .network_models.models.PointNet2_tcls import PN2_CLS_CONFIG, PointNet2
        self._device = torch.device(device)
        if self._device.type != "cuda":
            raise RuntimeError("s4g_release_b200.GraspDetector needs a CUDA device (there is no CPU path)")
        self.model = model if model is not None else PointNet2(**PN2_CLS_CONFIG)
        if state_dict is not None:
            self.model.load_state_dict({k[7:] if k.startswith("module.") else k: v for k, v in state_dict.items()})
        self.model = self.model.to(self._device).eval()
        self.pre = GraspPreProcessor(num_input=num_input)
        self.post = GraspPostProcessor(camera2base=camera2base)

    def eval(self, cloud, rng=np.random):
        """grasp_detector.py:107-121: (3, n) cloud in the camera frame -> the four prediction tensors."""
        points = self.pre.pre_processing(torch.as_tensor(np.asarray(cloud, dtype=np.float32)).to(self._device), rng=rng)
        with torch.no_grad():
            return self.model({"scene_points": points.unsqueeze(0)})

    def detect(self, cloud_array, cloud_mask=None, num_selected=5, score_threshold=0.7, verticalness_threshold=0.2,
               collision_check=True, debug=False, rng=np.random, nms_min_dist=None):
        """grasp_detector.py:186-262.  ``rng``: numpy generator for the two random draws (sub-sample indices, then the
        sorted uniforms of the importance sampling — the order the reference consumes np.random in)."""
        cloud_array = np.asarray(cloud_array)
        assert cloud_array.ndim == 2, "Evaluation mode do not support batch, input should have shape (n, 3) or (3, n)."
        assert cloud_array.shape[0] == 3 or cloud_array.shape[1] == 3, \
            "input should have shape (n, 3) or (3, n), but given {}".format(cloud_array.shape)
        if cloud_array.shape[1] == 3:
            cloud_array = cloud_array.T
        target_cloud = cloud_array[:, cloud_mask] if isinstance(cloud_mask, np.ndarray) else cloud_array
        full = torch.as_tensor(np.ascontiguousarray(cloud_array, dtype=np.float32)).to(self._device)
        target = torch.as_tensor(np.ascontiguousarray(target_cloud, dtype=np.float32)).to(self._device)
        points = self.pre.pre_processing(target, rng=rng)
        with torch.no_grad():
            predictions = self.model({"scene_points": points.unsqueeze(0)})
        poses, scores = self.post.post_processing(points, predictions, score_threshold, verticalness_threshold, debug)
        if collision_check and poses.shape[0]:
            keep = self.post.collision_free(poses, full.t().contiguous())  # against the WHOLE input cloud (:219-221)
            poses, scores = poses[keep], scores[keep]
        if nms_min_dist is not None and poses.shape[0]:
            keep = self.post.nms(poses, scores, nms_min_dist)
            poses, scores = poses[keep], scores[keep]
        if poses.shape[0] > num_selected:
            pick = self.post.importance_sample(scores, np.sort(rng.rand(num_selected)))
            poses, scores = poses[pick], scores[pick]
        return poses.cpu().numpy(), scores.cpu().numpy()
