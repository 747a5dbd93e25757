"""A/B timing of LocalBundleAdjustment (configs[3]) across builds of the library: tools/ab_solve.py build/libcmos_A.so ..."""
import os
import subprocess
import sys

SNIPPET = r'''
import numpy as np
from ceres_mono_orb_slam2_b200 import CeresOptimizer, synth
from ceres_mono_orb_slam2_b200.ba_bench import _bench_graph
K4 = np.array(synth.KITTI_K, np.float32)
G = synth.make_ba_problem(20, 3000, 4, seed=4)
opt = CeresOptimizer(max_cams=20, max_points=3000, max_obs=12000)
r = _bench_graph(opt, G, K4, 20, 3, True, (5, 10), "x")
print("%.4f ms per solve, %.1f Mresid/s, cost %s" % (r["ms_per_solve"], r["value"], r["cost"][1]))
'''
for lib in sys.argv[1:]:
    env = dict(os.environ, CMOS_B200_LIB=os.path.abspath(lib))
    out = subprocess.run([sys.executable, "-c", SNIPPET], env=env, capture_output=True, text=True)
    print(lib, (out.stdout.strip().splitlines() or [out.stderr[-300:]])[-1])
