#!/usr/bin/env python
"""Progressive rendering through the C ABI — the viewer hook.

The reference's viewer calls `render` once and shows the finished frame (crates/viewer/src/render_output_view.rs:84-98).
With `rtcuda_render_samples_accumulate_device` a caller keeps ONE device plane of radiance sums, adds consecutive sample
ranges to it, and shows `plane / samples_so_far` after every pass: the image converges on screen, and after the last range
it is the frame `render` returns at that spp (same streams per sample; float sum associated per range).

  python examples/progressive_viewer.py [scene.npz] [spp] [out_prefix]    # writes out_prefix_<n>spp.npy per pass
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                      # only used as the owner of the device plane
import raytracing_cuda as rc

scene = rc.Scene.load_npz(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests/golden/scenes/cbbunny_area_light_transforms.npz"))
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
prefix = sys.argv[3] if len(sys.argv) > 3 else None
settings = rc.RaytracerSettings(samples_per_pixel=spp)
with rc.CudaRenderer(scene) as r:
    plane = torch.zeros((r.height, r.width, 3), dtype=torch.float32, device="cuda:0")
    torch.cuda.synchronize()
    lo, step = 0, 1
    while lo < spp:                                   # 1, 2, 4, 8, ... samples per pass: quick first image, long later passes
        hi = min(spp, lo + step)
        r.render_samples_accumulate_device(settings, lo, hi, plane.data_ptr())
        image = (plane / hi).cpu().numpy()            # <- what a viewer would blit
        print(f"{hi:5d} spp  mean {image.mean():.6f}  ({r.stats()['render_ms']:.1f} ms for samples [{lo}, {hi}))")
        if prefix:
            import numpy as np
            np.save(f"{prefix}_{hi}spp.npy", image)
        lo, step = hi, step * 2
