"""Does the 3-NN search of the last propagation level (25 600 queries x 5 120 keys per scene) overlap with the farthest
point sampling of set-abstraction level 1 (5 120 -> 1 024, one CTA per scene) when they run on two streams?
    python profiles/overlap_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s4g_release_b200.engine import FusedPointNet2 as E  # noqa: E402

B = 64
g = torch.Generator(device="cuda").manual_seed(0)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench import synthetic_scenes  # noqa: E402

xyz0 = synthetic_scenes(B, 1000).cuda()                    # the bench's tabletop clouds
xyz1 = E.gather_xyz(xyz0, E.fps(xyz0, 5120))               # level-1 centroids = farthest point samples
side = torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def both():
    ev = torch.cuda.Event()
    ev.record()
    E.fps(xyz1, 1024)
    side.wait_event(ev)
    with torch.cuda.stream(side):
        E.three_nn_weights(xyz0, xyz1)
    torch.cuda.current_stream().wait_stream(side)


print("fps level 1 alone      %.3f ms" % timed(lambda: E.fps(xyz1, 1024)))
print("3-NN level 0<-1 alone  %.3f ms" % timed(lambda: E.three_nn_weights(xyz0, xyz1)))
print("both, two streams      %.3f ms" % timed(both))
