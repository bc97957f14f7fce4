"""Quick per-stage timing of forward+backward on a synthetic scene (development aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from scenes import synthetic_scene, random_v_output
import brush_b200.render as R

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
h = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
shift = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0   # log-scale shift (config [3]: 4M Gaussians at 4K use -ln 2)
cam, tr, sh, op = synthetic_scene(n, w, h, scale_shift=shift)
ctx = R.RenderContext(n, w, h, int(sys.argv[5]) if len(sys.argv) > 5 else 0)
d = ctx.device
ttr, tsh, top = (torch.from_numpy(x).to(d) for x in (tr, sh, op))
vout = torch.from_numpy(random_v_output(h, w)).to(d)
def step():
    out = R.render_splats(ctx, cam, (w, h), ttr, tsh, top)
    vc = R.rasterize_bwd(out, vout)
    g = R.project_bwd(out, ttr, tsh, top, vc)
    return out, g
for _ in range(3): out, g = step()
torch.cuda.synchronize()
out = R.render_splats(ctx, cam, (w, h), ttr, tsh, top)
print("blend stats:", R.blend_stats(out, vout))
print("V", out.num_visible, "I", out.num_intersections, "overflow", out.intersection_overflow, "arena MB", ctx.arena_bytes() / 1e6)
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
tf = tb = tp = 0.0
K = 10
for _ in range(K):
    e[0].record(); out = R.render_splats(ctx, cam, (w, h), ttr, tsh, top); e[1].record()
    vc = R.rasterize_bwd(out, vout); e[2].record()
    g = R.project_bwd(out, ttr, tsh, top, vc); e[3].record()
    torch.cuda.synchronize()
    tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2]); tp += e[2].elapsed_time(e[3])
print(f"forward {tf/K:.3f} ms  raster_bwd {tb/K:.3f} ms  project_bwd {tp/K:.3f} ms  total {(tf+tb+tp)/K:.3f} ms  -> {w*h/((tf+tb+tp)/K*1e-3)/1e6:.1f} Mpix/s")
