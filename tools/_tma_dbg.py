import numpy as np
from ceres_mono_orb_slam2_b200 import ORBextractor, synth
img = synth.make_image(1241, 376, 11)
ext = ORBextractor(2000, 1.2, 8, 20, 7, max_width=1241, max_height=376, max_batch=1)
kps, desc = ext(img)
print(len(kps))
