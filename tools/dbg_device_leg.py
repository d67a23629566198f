"""Development aid: one camera stream of the bench, device-resident leg, per-frame stage times."""
import sys, time, types
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from openvslam_b200 import feature, _lib

cfg = bench.CONFIGS[4]
dev = torch.device("cuda", 0)
W, H, NKP = cfg["W"], cfg["H"], cfg["NKP"]
ring = 12
wl = bench.make_workload(cfg, 0, ring)
h = torch.empty((ring, H, W), dtype=torch.uint8).pin_memory()
for i, f in enumerate(wl["frames"]):
    h[i].copy_(torch.from_numpy(f))
d = h.to(dev)
ext0 = feature.orb_extractor(feature.orb_params(max_num_keypts=NKP), device=0)
lmsets = bench.make_landmark_sets(cfg, wl, ext0)
ext0.close()
nstreams = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cams = [bench.CameraStream(cfg, s, 0, dev, d, h.numpy(), None, None, wl, lmsets, ring, 4, 8 if nstreams == 1 else 2) for s in range(nstreams)]
import threading
def work(cs, leg):
    torch.cuda.set_device(0)
    for i in range(8):
        t0 = time.perf_counter()
        n = getattr(cs, leg)(i)
        if cs.sid == 0:
            print(leg, i, "n", n if leg == "step_device" else "-", "ms %.3f" % ((time.perf_counter() - t0) * 1e3),
                  "requeries", cs.mt.num_requeries() if hasattr(cs.mt, "num_requeries") else "?",
                  {k: round(v, 1) for k, v in cs.ext.last_timings_us().items()}, flush=True)
for leg in ("step_device", "step_host", "step_device"):
    ths = [threading.Thread(target=work, args=(cs, leg)) for cs in cams]
    [t.start() for t in ths]; [t.join() for t in ths]
    torch.cuda.synchronize()
    print({k: (cams[0].stage_ms["device"] / max(1, cams[0].st_dev["frames"])).round(3).tolist() for k in ["dev"]})
    for cs in cams: cs.reset()
