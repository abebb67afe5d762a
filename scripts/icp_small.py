import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from followmyhold_b200.alignment.mesh_align import icp_points, icp_points_many
rng = np.random.default_rng(0)
tgt = rng.normal(size=(3000, 3)); src = (tgt[:2000] - 0.03) / 1.04 + 0.004 * rng.normal(size=(2000, 3))
print(icp_points(src, tgt, 4, 400, False, 0.7, 3.0)[1])
print([c for _, c in icp_points_many([(src, tgt), (src[:1500], tgt)], 3, [400, 300], False, 0.7, 3.0)])
