"""A/B of render-kernel variants on the bench frame (resident scene + tree): kernel ms, counters, frame identity.
usage: python tools/ab_render.py option=a,b[,c] [shadows]   e.g. hull=0,1   (GPU box; option = an rtds_set_option name)"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

rt = entry.load_rtds()
var, vals = sys.argv[1].split("=")
vals = vals.split(",")
shadows = 1 if len(sys.argv) > 2 and sys.argv[2] == "shadows" else 0
v = np.fromfile(os.path.join(ROOT, "tests", "golden", "bunny_vertices.f32"), np.float32).reshape(-1, 3)
clones = int(os.environ.get("AB_CLONES", "30"))
W, H, SPP = int(os.environ.get("AB_W", "3840")), int(os.environ.get("AB_H", "2160")), int(os.environ.get("AB_SPP", "4"))
sph, mat = rt.scene_from_vertices(v, clones)
ctx = rt.Rtds(0)
ctx.set_spheres(sph, mat)
ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
out = np.zeros((H, W, 3), np.uint8)
dev = None
if os.environ.get("AB_DEVICE"):            # resident path: rtds_render_device into a device buffer (no download, no bands)
    import torch
    dev = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
ref = None
for rep in range(2):
    for val in vals:
        ctx.set_option(var.lower().replace("rtds_", ""), int(val))
        ms = []
        for i in range(6):
            if dev is not None:
                st = ctx.render_device(rt.LBVH, ctx.render_params(W, H, SPP, shadows=shadows), dev.data_ptr())
                rgb = dev.cpu().numpy()
            else:
                rgb, _, _, st = ctx.render(rt.LBVH, W, H, SPP, out=out, shadows=shadows)
            ms.append(st["ms_kernel"])
        md5 = hashlib.md5(rgb.tobytes()).hexdigest()
        if ref is None:
            ref = md5
        print("%s=%s: kernel %.3f ms (min %.3f) | per ray: %.2f slab tests, %.2f node visits, %.3f prim tests | frame %s" %
              (var, val, float(np.median(ms)), min(ms), st["node_tests"] / st["rays"], st["node_visits"] / st["rays"],
               st["prim_tests"] / st["rays"], "identical" if md5 == ref else "DIFFERENT " + md5))
