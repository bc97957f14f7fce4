"""View-sharded data parallelism (no counterpart in the single-device reference).

Every rank holds the full parameter set and renders its own view(s) of the step's batch; one exchange per step makes
the gradient of the mean-over-views loss (train.rs:254-260) and the refine statistics (stats.rs:40-50: MAX / SUM / MAX)
available on every rank.  Parity definition: equal to the single-GPU step that accumulates the views'
gradients sequentially, up to f32 summation order.

Two layers live here:
  * DpComm -- the product path: one NCCL rank of the LIBRARY's communicator (csrc/dp.cu).  bg_train_step_views and
    bg_dp_exchange run the SH-factored exchange (all-reduce SUM 48 N B + all-reduce MAX 8 N B + all-gather 12 local N B
    per rank) on the library's own stream; torch.distributed only carries the 128-byte NCCL id.
  * FlatGradients / FactoredGradients / ShFactoredReducer / ViewShardedReducer -- the same exchanges written over
    torch.distributed collectives for hosts that drive the operators themselves (SplatTrainer.grad_hook) and for the
    world_size-2 gloo tests that pin the host-side logic without a GPU (tests/test_dp_gloo.py).  ViewShardedReducer is the
    plain form SURVEY 8e describes: ONE all-reduce over a flat (10+3K+1) N float buffer.
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.distributed as dist


class DpComm:
    """One NCCL rank of the library's own communicator (bg_dp_comm_create), bound to the render context's device.
    The 128-byte NCCL id is made by rank 0 and broadcast through torch.distributed (whatever backend it runs on);
    everything after that -- the gradient exchange of bg_train_step_views / bg_dp_exchange -- happens inside
    libbrush_b200.so on its own stream.  A Rust host would pass the id through its own rendezvous."""

    def __init__(self, ctx, group=None):
        import ctypes as C
        from . import _lib
        if not (dist.is_initialized() and dist.get_world_size(group) > 1):
            raise RuntimeError("DpComm needs an initialised torch.distributed group with more than one rank")
        lib = _lib.load()
        self.rank, self.world, self.ctx = dist.get_rank(group), dist.get_world_size(group), ctx
        ident = (C.c_uint8 * 128)()
        if self.rank == 0:
            _lib.check(lib.bg_dp_unique_id(ident), "bg_dp_unique_id")
        on_gpu = dist.get_backend(group) == "nccl"
        t = torch.tensor(list(ident), dtype=torch.uint8, device=ctx.device if on_gpu else "cpu")
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = (C.c_uint8 * 128)(*t.cpu().tolist())
        h = C.c_void_p()
        _lib.check(lib.bg_dp_comm_create(ctx.handle, ident, self.rank, self.world, C.byref(h)), "bg_dp_comm_create")
        self.handle = h

    def pack_view(self, n: int, local_views: int, view: int, first: bool, v_t, v_o, v_color, v_refine, visible, max_radius,
                  small: torch.Tensor, stat: torch.Tensor, record: torch.Tensor):
        """bg_dp_pack_view: fold one view's operator outputs into the interleaved exchange rows (csrc/bg_dp.cuh)."""
        from . import _lib
        from .render import _stream_ptr
        _lib.check(_lib.load().bg_dp_pack_view(self.ctx.handle, _stream_ptr(self.ctx.device), n, local_views, view, int(first),
                                               v_t.data_ptr(), v_o.data_ptr(), v_color.data_ptr(), v_refine.data_ptr(),
                                               visible.data_ptr(), max_radius.data_ptr(), small.data_ptr(), stat.data_ptr(),
                                               record.data_ptr()),
                   "bg_dp_pack_view")

    def exchange(self, n: int, local_views: int, small: torch.Tensor, stat: torch.Tensor, record: torch.Tensor, recv: torch.Tensor,
                 chunks: int = 1):
        """bg_dp_exchange: all-reduce `small` (SUM) and `stat` (MAX) in place, all-gather `record` into `recv` (csrc/bg_dp.cuh)."""
        from . import _lib
        from .render import _stream_ptr
        _lib.check(_lib.load().bg_dp_exchange(self.ctx.handle, self.handle, _stream_ptr(self.ctx.device), n, local_views,
                                              small.data_ptr(), stat.data_ptr(), record.data_ptr(), recv.data_ptr(), chunks),
                   "bg_dp_exchange")

    def close(self):
        if self.handle:
            from . import _lib
            _lib.load().bg_dp_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FlatGradients:
    """One allocation holding v_transforms [n,10] | v_sh [n,k,3] | v_raw_opac [n]: project_bwd writes
    straight into the views, the all-reduce runs once over `flat` with no packing copies."""

    def __init__(self, n: int, k: int, device):
        self.n, self.k = n, k
        up4 = lambda x: (x + 3) // 4 * 4                # every segment starts 16-byte aligned (128-bit row access)
        o0 = up4(n * 10)
        o1 = o0 + up4(n * 3 * k)
        self.flat = torch.zeros(o1 + n, dtype=torch.float32, device=device)
        self.v_t = self.flat[:n * 10].view(n, 10)
        self.v_sh = self.flat[o0:o0 + n * 3 * k].view(n, k, 3)
        self.v_o = self.flat[o1:o1 + n].view(n)
        self.v_r = torch.empty(n, dtype=torch.float32, device=device)  # refine weight: MAX-reduced separately

    def outputs(self):
        return self.v_t, self.v_sh, self.v_o, self.v_r


class FactoredGradients:
    """Exchange buffers of the SH-factored scheme, two collectives per step:
      `small`  = v_transforms [n,10] | v_raw_opac [n] | visible [n]           -> ONE all-reduce (SUM), 48 B/Gaussian
      `record` = v_color [n,3] | v_refine [n] | max_radius [n] of this view   -> ONE all-gather, 20 B/Gaussian/view
    v_sh [n,k,3] is rebuilt locally from the gathered colours; the refine statistics that need MAX (stats.rs:40-50)
    are reduced locally over the gathered records.  A rank sends 68 n bytes instead of (44 + 12 k) n + 12 n."""

    def __init__(self, n: int, k: int, views: int, device):
        self.n, self.k, self.views = n, k, views
        self.small = torch.zeros(n * 12, dtype=torch.float32, device=device)
        self.v_t = self.small[:n * 10].view(n, 10)
        self.v_o = self.small[n * 10:n * 11].view(n)
        self.visible = self.small[n * 11:].view(n)
        self.record = torch.zeros(n * 5, dtype=torch.float32, device=device)
        self.v_color = self.record[:n * 3].view(n, 3)
        self.v_r = self.record[n * 3:n * 4].view(n)
        self.max_radius = self.record[n * 4:].view(n)
        self.records_all = torch.empty((views, n * 5), dtype=torch.float32, device=device)
        self.v_sh = torch.empty((n, k, 3), dtype=torch.float32, device=device)

    def outputs(self):
        """For render.project_bwd_factored(outputs=...)."""
        return self.v_t, self.v_color, self.v_o, self.v_r

    def gradients(self):
        return self.v_t, self.v_sh, self.v_o, self.v_r


class ShFactoredReducer:
    """One view per rank.  all-reduce(small) + all-gather(record) + local rebuild of v_sh in view order, so every
    rank ends with bit-identical gradients (a property a ring all-reduce also has, but here it follows from the
    fixed summation order of bg_sh_grad_from_views)."""

    def __init__(self, ctx, num_views_total: int, group=None):
        self.ctx, self.views, self.group = ctx, num_views_total, group

    def reduce(self, fg: FactoredGradients, transforms: torch.Tensor, cam_positions, visible: torch.Tensor = None,
               max_radius: torch.Tensor = None) -> None:
        """In place on fg (gradients) and, when given, on `visible` (SUM) / `max_radius` (MAX); fg.v_r gets MAX."""
        from .render import sh_grad_from_views
        n = fg.n
        multi = dist.is_initialized() and dist.get_world_size(self.group) > 1
        if visible is not None:
            fg.visible.copy_(visible)
        if max_radius is not None:
            fg.max_radius.copy_(max_radius)
        if multi:
            dist.all_gather_into_tensor(fg.records_all.view(-1), fg.record, group=self.group)
            dist.all_reduce(fg.small, op=dist.ReduceOp.SUM, group=self.group)
        else:
            fg.records_all[0].copy_(fg.record)
        inv = 1.0 / self.views
        sh_grad_from_views(self.ctx, transforms, fg.k, cam_positions, fg.records_all, inv, out=fg.v_sh, view_stride=5 * n)
        if self.views != 1:
            fg.small[:11 * n].mul_(inv)
        torch.amax(fg.records_all[:, 3 * n:4 * n], dim=0, out=fg.v_r)
        if visible is not None:
            visible.copy_(fg.visible)
        if max_radius is not None:
            torch.amax(fg.records_all[:, 4 * n:], dim=0, out=max_radius)


class ViewShardedReducer:
    def __init__(self, num_views_total: int, group=None):
        self.views = num_views_total
        self.group = group
        self._flat = None

    def reduce_flat(self, fg: FlatGradients) -> None:
        """In place on fg.flat: sum over ranks / views, ONE collective, no copies."""
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(fg.flat, op=dist.ReduceOp.SUM, group=self.group)
        if self.views != 1:
            fg.flat.mul_(1.0 / self.views)

    def _flat_buf(self, tensors: Sequence[torch.Tensor]) -> torch.Tensor:
        total = sum(t.numel() for t in tensors)
        if self._flat is None or self._flat.numel() != total or self._flat.device != tensors[0].device:
            self._flat = torch.empty(total, dtype=torch.float32, device=tensors[0].device)
        return self._flat

    def reduce_gradients(self, grads: Sequence[torch.Tensor]) -> None:
        """In place: grads <- sum over ranks / views.  One collective for all tensors."""
        flat = self._flat_buf(grads)
        off = 0
        for g in grads:
            flat[off:off + g.numel()].copy_(g.reshape(-1))
            off += g.numel()
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.mul_(1.0 / self.views)
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    def reduce_stats(self, refine_weight: torch.Tensor, visible: torch.Tensor, max_radius: torch.Tensor) -> None:
        """stats.rs:40-50 across ranks: MAX(refine), SUM(visible), MAX(max_radius).  Two collectives."""
        if not (dist.is_initialized() and dist.get_world_size(self.group) > 1):
            return
        mx = torch.stack([refine_weight, max_radius])
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)
        refine_weight.copy_(mx[0])
        max_radius.copy_(mx[1])
        dist.all_reduce(visible, op=dist.ReduceOp.SUM, group=self.group)

    def hook(self, tensors: Sequence[torch.Tensor]) -> None:
        """SplatTrainer.grad_hook signature: (v_t, v_sh, v_o, v_refine, visible, max_radius)."""
        v_t, v_sh, v_o, v_r, visible, max_radius = tensors
        self.reduce_gradients((v_t, v_sh, v_o))
        self.reduce_stats(v_r, visible, max_radius)
