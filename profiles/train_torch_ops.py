"""Which torch (library / element-wise) ops are left in one fused training step, with their input shapes:
    python profiles/train_torch_ops.py > gpurun_out/train_torch_ops.txt
torch.profiler, CUDA activities, grouped by (op, input shapes), sorted by device time.  Not a timing source for the
bench (profiler overhead) — it tells WHERE the at:: kernels of the ncu launch list come from."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from bench import NUM_POINTS, synthetic_scenes  # noqa: E402
from s4g_release_b200.network_models.models.PointNet2_tcls import PN2_CLS_CONFIG, PointNet2, PointNet2Loss  # noqa: E402
from s4g_release_b200.train import Trainer, synthetic_labels  # noqa: E402

B = int(os.environ.get("S4G_PROFILE_BATCH", "32"))
torch.manual_seed(0)
model = PointNet2(**PN2_CLS_CONFIG).cuda()
trainer = Trainer(model, PointNet2Loss(), fused=True)
x = synthetic_scenes(min(B, 8), 1000).repeat((B + 7) // 8, 1, 1)[:B].contiguous().cuda()
y = synthetic_labels(B, NUM_POINTS, 4000, 2000, device="cuda")
for _ in range(2):
    trainer.step({"scene_points": x}, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    trainer.step({"scene_points": x}, y)
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=60, max_name_column_width=48,
                                                         max_shapes_column_width=70))
