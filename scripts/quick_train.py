"""Per-kernel view of one SplatTrainer.step (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from scenes import synthetic_scene
import brush_b200.render as R, brush_b200.train as T
n, w, h = 1_000_000, 1920, 1080
cam, tr, sh, op = synthetic_scene(n, w, h)
ctx = R.RenderContext(n, w, h, 0)
d = ctx.device
splats = T.Splats(*(torch.from_numpy(x).to(d) for x in (tr, sh, op)))
gt = torch.randint(0, 2 ** 31 - 1, (h, w), dtype=torch.int32, device=d) | (255 << 24)
trainer = T.SplatTrainer(T.TrainConfig(), ctx, T.bounds_from_pos(0.8, tr[:, :3]))
batch = T.SceneBatch(img_packed=gt, camera=cam)
for _ in range(5):
    trainer.step(batch, splats)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 20
e0.record()
for _ in range(K):
    trainer.step(batch, splats)
e1.record(); torch.cuda.synchronize()
print(f"train step {e0.elapsed_time(e1)/K:.3f} ms -> {1e3*K/e0.elapsed_time(e1):.1f} it/s")
