import numpy as np, sys
sys.path.insert(0, '.')
from openvslam_b200 import feature
rng = np.random.default_rng(1)
img = (rng.random((480, 640)) * 255).astype(np.uint8)
ex = feature.orb_extractor(feature.orb_params(max_num_keypts=1000))
kp, desc = ex.extract(img)
print("ok", len(kp))
