"""Device-side grasp post-processing — the tail of the reference's ``GraspDetector`` (grasp_detector.py:124-251)
on the sm_100a kernels of csrc/postprocess.cu, batched over scenes.

    post = GraspPostProcessor()
    poses, scores = post.post_processing(points, predictions, 0.7, 0.2)      # reference signature, scene 0
    res = post.detect_batch(points, predictions, clouds=..., num_selected=5) # all scenes of the batch

Constants restate configs/real_world_config.py:21-24, grasp_detector.py:26-27,177, configs/gripper_config.py:9-21
and configs/processing_config.py:19,37-40.  There is no CPU path: every entry point needs CUDA tensors.
"""
import ctypes

import numpy as np
import torch

from ._lib import check, lib, ptr, stream_ptr

REAL2TRAIN = np.array([[0, 1, 0, 0], [1, 0, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]], dtype=np.float64)
TRAIN2REAL = np.linalg.inv(REAL2TRAIN)
CAMERA2BASE = np.array([[-0.00377177, 0.54720216, -0.83699198, 0.766],
                        [0.99981506, -0.01372054, -0.01347562, -0.276],
                        [-0.01885787, -0.83688801, -0.54704921, 0.62],
                        [0., 0., 0., 1.]])
T_SCORE = np.array([0.08, 0.06, 0.04, 0.02], dtype=np.float64)
HALF_BOTTOM_WIDTH, BOTTOM_LENGTH, FINGER_WIDTH, HALF_HAND_THICKNESS, FINGER_LENGTH = 0.057, 0.16, 0.023, 0.012, 0.09
HALF_BOTTOM_SPACE = HALF_BOTTOM_WIDTH - FINGER_WIDTH
BACK_COLLISION_MARGIN, BACK_COLLISION_THRESHOLD, FINGER_COLLISION_THRESHOLD = 0.0, 10 * np.sqrt(8), 10


def _dptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class GraspPostProcessor:
    def __init__(self, camera2base=CAMERA2BASE, max_candidates=4096):
        self.camera2base = np.asarray(camera2base, dtype=np.float64)
        # x_direction = -camera2base_R @ TRAIN2REAL_R @ rotation[:, :, 0]; only its z component is used (:154-156)
        self._approach_row = np.ascontiguousarray((-self.camera2base[:3, :3] @ TRAIN2REAL[:3, :3])[2])
        self._train2real = np.ascontiguousarray(TRAIN2REAL.reshape(-1))
        self._gripper = np.array([FINGER_LENGTH, BOTTOM_LENGTH, HALF_HAND_THICKNESS, HALF_BOTTOM_WIDTH, HALF_BOTTOM_SPACE,
                                  BACK_COLLISION_MARGIN, BACK_COLLISION_THRESHOLD, FINGER_COLLISION_THRESHOLD],
                                 dtype=np.float32)
        self.max_candidates = int(max_candidates)

    # ------------------------------------------------------------------ batched core
    @staticmethod
    def _need_cuda(*tensors):
        for t in tensors:
            if not t.is_cuda:
                raise RuntimeError("grasp post-processing needs CUDA tensors (there is no CPU path)")

    def scores(self, score_logits):
        """(B, C, N) fp32 logits -> (B, N) fp64 expected grasp score (grasp_detector.py:142-145)."""
        self._need_cuda(score_logits)
        x = score_logits.float().contiguous()
        B, C, N = x.shape
        out = torch.empty((B, N), dtype=torch.float64, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.s4g_grasp_scores_f32(ptr(x), B, C, N, ptr(out), stream_ptr(x.device)), "grasp_scores")
        return out

    def select_and_decode(self, points, predictions, score_threshold=0.7, vertical_degree_threshold=0.2,
                          check_overflow=True):
        """All scenes at once.  points (B,3,N) fp32.  Returns dict with ``n`` (B,) int32 candidate counts,
        ``poses`` (B, cap, 4, 4) fp64, ``scores`` (B, cap) fp64, ``point_index`` (the cloud point each pose is anchored at) / ``rotation_index`` (the rank whose rotation block
        the reference's reshape pairs with it) (B, cap) int32,
        ``all_scores`` (B, N) fp64; rows past ``n[b]`` are undefined."""
        pts = points.float().contiguous()
        fr = predictions["frame_R"].float().contiguous()
        ft = predictions["frame_t"].float().contiguous()
        self._need_cuda(pts, fr, ft)
        B, _, N = pts.shape
        dev = pts.device
        all_scores = self.scores(predictions["score"])
        work = torch.empty(2 * B * N, dtype=torch.int32, device=dev)
        n_out = torch.empty(B, dtype=torch.int32, device=dev)
        n_high = torch.empty(B, dtype=torch.int32, device=dev)
        cap = min(self.max_candidates, N)
        with torch.cuda.device(dev):
            while True:
                out_point = torch.empty((B, cap), dtype=torch.int32, device=dev)
                out_rot = torch.empty((B, cap), dtype=torch.int32, device=dev)
                check(lib.s4g_grasp_select(ptr(all_scores), ptr(fr), B, N, float(score_threshold),
                                           float(vertical_degree_threshold), _dptr(self._approach_row), ptr(work), cap,
                                           ptr(out_point), ptr(out_rot), ptr(n_out), ptr(n_high), stream_ptr(dev)),
                      "grasp_select")
                if not check_overflow:  # streaming use: no host round trip; candidates past `cap` are dropped
                    break
                need = int(n_out.max().item()) if B else 0
                if need <= cap:
                    break
                cap = min(N, max(need, 2 * cap))  # rare: more candidates than the default capacity
            poses = torch.empty((B, cap, 4, 4), dtype=torch.float64, device=dev)
            scores = torch.empty((B, cap), dtype=torch.float64, device=dev)
            check(lib.s4g_grasp_poses(ptr(pts), ptr(fr), ptr(ft), ptr(all_scores), B, N, ft.shape[1], ptr(out_point),
                                      ptr(out_rot), ptr(n_out), ptr(n_high), ptr(work), cap, _dptr(self._train2real),
                                      _dptr(T_SCORE), ptr(poses), ptr(scores), stream_ptr(dev)), "grasp_poses")
        return {"n": n_out, "n_high": n_high, "poses": poses, "scores": scores, "point_index": out_point,
                "rotation_index": out_rot, "all_scores": all_scores}

    def collision_free(self, poses, cloud_n3, return_counts=False):
        """poses (n,4,4) fp64, cloud (m,3) fp32 -> bool mask (n,) of collision-free grasps
        (view_collision_checker.py:37-65 for every pose)."""
        self._need_cuda(poses, cloud_n3)
        p = poses.double().contiguous()
        c = cloud_n3.float().contiguous()
        n = p.shape[0]
        ok = torch.empty(n, dtype=torch.uint8, device=p.device)
        counts = torch.empty((n, 2), dtype=torch.int32, device=p.device)
        with torch.cuda.device(p.device):
            check(lib.s4g_grasp_collision_f32(ptr(p), n, ptr(c), c.shape[0], _dptr(self._gripper), ptr(ok), ptr(counts),
                                              stream_ptr(p.device)), "grasp_collision")
        return (ok.bool(), counts) if return_counts else ok.bool()

    def nms(self, poses, scores, min_dist):
        """Greedy translation de-duplication in descending score order (ties: lower index first).  OUR definition —
        the reference ships none (README.md:58; sketch at utils/file_logger_cls.py:220-225).  Returns kept indices."""
        self._need_cuda(poses, scores)
        p = poses.double().contiguous()
        n = p.shape[0]
        if n == 0:
            return torch.empty(0, dtype=torch.int64, device=p.device)
        order = torch.sort(scores.double(), descending=True, stable=True)[1].int().contiguous()
        kept = torch.empty(n, dtype=torch.int32, device=p.device)
        n_kept = torch.empty(1, dtype=torch.int32, device=p.device)
        with torch.cuda.device(p.device):
            check(lib.s4g_grasp_nms(ptr(p), ptr(order), n, float(min_dist), ptr(kept), ptr(n_kept), stream_ptr(p.device)),
                  "grasp_nms")
        return kept[: int(n_kept.item())].long()

    def importance_sample(self, scores, sorted_uniform):
        """grasp_detector.py:235-246; ``sorted_uniform`` = np.sort(np.random.rand(k)) (host array or tensor)."""
        self._need_cuda(scores)
        s = scores.double().contiguous()
        u = torch.as_tensor(np.asarray(sorted_uniform, dtype=np.float64), device=s.device)
        cum = torch.empty_like(s)
        picked = torch.empty(u.numel(), dtype=torch.int32, device=s.device)
        with torch.cuda.device(s.device):
            check(lib.s4g_grasp_importance_sample(ptr(s), s.numel(), ptr(u), u.numel(), ptr(cum), ptr(picked),
                                                  stream_ptr(s.device)), "grasp_importance_sample")
        return picked.long()

    # ------------------------------------------------------------------ reference-shaped entry points
    def post_processing(self, points_array, predictions, score_threshold, vertical_degree_threshold, debug=False):
        """GraspDetector.post_processing (grasp_detector.py:137-185): scene 0 of the batch ->
        (poses (n,4,4) fp64, scores (n,) fp64), both CUDA tensors."""
        pts = points_array if torch.is_tensor(points_array) else torch.as_tensor(np.asarray(points_array))
        pts = pts.to(predictions["score"].device)
        if pts.dim() == 2:
            if pts.shape[0] != 3:
                pts = pts.t()
            pts = pts.unsqueeze(0)
        r = self.select_and_decode(pts[:1], {k: v[:1] for k, v in predictions.items()}, score_threshold,
                                   vertical_degree_threshold)
        n = int(r["n"][0].item())
        return r["poses"][0, :n], r["scores"][0, :n]

    def detect_batch_device(self, points, predictions, clouds=None, num_selected=5, score_threshold=0.7,
                            verticalness_threshold=0.2, collision_check=True, nms_min_dist=None, sorted_uniform=None):
        """``detect_batch`` without any host round trip (BASELINE config 5: thousands of scenes streamed through a
        GPU): five kernel launches for the whole batch.  ``clouds``: (B,3,m) fp32 CUDA tensor for the collision
        test (default: the network input); ``sorted_uniform``: (B, num_selected) ascending uniforms per scene
        (np.sort(np.random.rand(k)) in the reference, grasp_detector.py:236) or None to keep the first
        ``num_selected``.  Scenes with more than ``max_candidates`` candidates are truncated to the first ones.
        Returns dict of CUDA tensors: ``n`` (B,), ``poses`` (B,k,4,4), ``scores`` (B,k), ``index`` (B,k), ``n_candidates`` (B,)."""
        r = self.select_and_decode(points, predictions, score_threshold, verticalness_threshold, check_overflow=False)
        poses, scores, n_cand = r["poses"], r["scores"], r["n"]
        B, cap = scores.shape
        dev = poses.device
        m = int(num_selected)
        cloud = None
        if collision_check:
            cloud = (clouds if clouds is not None else points).float().contiguous()
            self._need_cuda(cloud)
        u = None
        if sorted_uniform is not None:
            u = torch.as_tensor(np.asarray(sorted_uniform, dtype=np.float64) if not torch.is_tensor(sorted_uniform)
                                else sorted_uniform, dtype=torch.float64).to(dev).contiguous()
            assert u.shape == (B, m)
        ws_bytes = int(lib.s4g_grasp_finish_batch_workspace(B, cap))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        out_index = torch.empty((B, m), dtype=torch.int32, device=dev)
        out_n = torch.empty(B, dtype=torch.int32, device=dev)
        out_poses = torch.empty((B, m, 4, 4), dtype=torch.float64, device=dev)
        out_scores = torch.empty((B, m), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            check(lib.s4g_grasp_finish_batch(ptr(poses), ptr(scores), ptr(n_cand), B, cap,
                                             ptr(cloud) if cloud is not None else None,
                                             cloud.shape[2] if cloud is not None else 0, _dptr(self._gripper),
                                             float(nms_min_dist) if nms_min_dist else 0.0,
                                             ptr(u) if u is not None else None, m, ptr(ws), ws_bytes, ptr(out_index),
                                             ptr(out_n), ptr(out_poses), ptr(out_scores), stream_ptr(dev)),
                  "grasp_finish_batch")
        return {"n": out_n, "poses": out_poses, "scores": out_scores, "index": out_index, "n_candidates": n_cand}

    def detect_batch(self, points, predictions, clouds=None, num_selected=5, score_threshold=0.7,
                     verticalness_threshold=0.2, collision_check=True, nms_min_dist=None, rng=None):
        """The part of GraspDetector.detect after the forward (grasp_detector.py:212-251) for every scene of the
        batch.  ``clouds``: list of (m,3) CUDA tensors for the collision check (default: the network input).
        Returns a list of (poses, scores) CUDA tensors."""
        r = self.select_and_decode(points, predictions, score_threshold, verticalness_threshold)
        counts = r["n"].tolist()
        rng = rng or np.random
        out = []
        for b, n in enumerate(counts):
            poses, scores = r["poses"][b, :n], r["scores"][b, :n]
            if collision_check and n:
                cloud = clouds[b] if clouds is not None else points[b].t()
                keep = self.collision_free(poses, cloud)
                poses, scores = poses[keep], scores[keep]
            if nms_min_dist is not None and poses.shape[0]:
                keep = self.nms(poses, scores, nms_min_dist)
                poses, scores = poses[keep], scores[keep]
            if poses.shape[0] > num_selected:
                pick = self.importance_sample(scores, np.sort(rng.rand(num_selected)))
                poses, scores = poses[pick], scores[pick]
            out.append((poses, scores))
        return out
