"""One fused training step at the BASELINE config[3] shape (32 scenes x 25 600 points) for ncu launch lists:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train.csv \
        python profiles/one_train_step.py ;  python profiles/one_train_step.py --summarize gpurun_out/train.csv"""
import csv
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 2 and sys.argv[1] == "--summarize":
    rows = [r for r in csv.reader(open(sys.argv[2], errors="replace")) if len(r) > 10]
    head = rows[0]
    ix = {n: head.index(n) for n in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    tot = {}
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).strip()
        name = re.sub(r"<.*", "", name)[:60]
        v = float(r[ix["Metric Value"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ix["Metric Unit"]], 1e-6)
        t = tot.setdefault(name, [0.0, 0])
        t[0] += v
        t[1] += 1
    total = sum(v[0] for v in tot.values())
    print("total %.2f ms over %d launches" % (total, sum(v[1] for v in tot.values())))
    for name, (ms, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:40]:
        print("%9.3f ms %5.1f%% %5d x  %s" % (ms, 100 * ms / total, n, name))
    sys.exit(0)

import torch  # noqa: E402

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
torch.cuda.profiler.start()
trainer.step({"scene_points": x}, y)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
