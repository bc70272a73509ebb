"""Golden vectors for the pre-processing step, produced by the REFERENCE's own code (run in the build container, where
/root/reference exists): utils/math_utils.py::transform_numpy_points is cut out of the source file and executed, and
GraspDetector.sample_single_cloud's index draw (grasp_detector.py:82-91) is replayed with a seeded numpy generator
(grasp_detector.py itself cannot be imported here: open3d / yacs are absent).  open3d's voxel / outlier calls are
no-ops in _pre_processing (their return values are dropped, cloud_processor.py:31-42), so the output below IS the
reference's network input.   python tests/golden/make_preprocess_golden.py -> tests/golden/preprocess_ref.npz"""
import os
import re

import numpy as np

REF = "/root/reference/inference/grasp_proposal"
HERE = os.path.dirname(os.path.abspath(__file__))

src = open(os.path.join(REF, "utils/math_utils.py")).read()
ns = {}
exec("import numpy as np\n" + re.search(r"def transform_numpy_points.*?return cloud_array\[:3, :\]\n", src, re.S).group(0), ns)
det = open(os.path.join(REF, "grasp_detector.py")).read()
real2train = eval(re.search(r"_REAL2TRAIN = (np\.array\(.*?\]\]\))", det, re.S).group(1))

out = {}
for name, n, m in (("large", 6000, 2048), ("small", 700, 2048)):
    rs = np.random.RandomState(len(name))
    cloud = (rs.rand(3, n) - 0.5).astype(np.float32)
    np.random.seed(17)  # sample_single_cloud uses the global generator (grasp_detector.py:86-89)
    if cloud.shape[1] > m:
        index = np.random.choice(np.arange(cloud.shape[1]), m, replace=False)
    else:
        index = np.random.choice(np.arange(cloud.shape[1]), m, replace=True)
    points = ns["transform_numpy_points"](cloud, real2train)[:, index]
    out[name + "/cloud"], out[name + "/index"] = cloud, index
    out[name + "/points_f32"] = points.astype(np.float32)  # torch.tensor(points, dtype=torch.float32) at :113
np.savez_compressed(os.path.join(HERE, "preprocess_ref.npz"), **out)
print({k: v.shape for k, v in out.items()})
