"""Test-side restatement of SplatTrainer::refine (brush-train/src/train.rs:431-893, multinomial.rs, quat_vec.rs) with
generic torch ops on whatever device the tensors live on -- statement by statement what the reference does with burn
tensors.  The product path is csrc/refine.cu (bg_refine); this file is what its tests compare against, and what the
CPU property tests (tests/test_refine_cpu.py) exercise.  Not imported by anything under brush_b200/."""
from typing import Optional

import numpy as np
import torch

from brush_b200.train import (BOUND_PERCENTILE, FRAC_1_SQRT_2, MIN_OPACITY, BoundingBox, RefineStats)


def multinomial_sample(weights: torch.Tensor, n: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Weighted sampling WITHOUT replacement (multinomial.rs:1-26; rand's `sample_weighted` is the
    Efraimidis-Spirakis scheme).  Non-finite or negative weights count as zero; at most
    #(positive weights) indices are returned (all-zero weights -> empty), no duplicates.
    Runs on the weights' device; a seeded generator makes data-parallel ranks agree."""
    w = torch.where(torch.isfinite(weights) & (weights >= 0), weights, torch.zeros_like(weights)).double()
    positive = int((w > 0).sum().item())
    n = min(int(n), positive)
    if n <= 0:
        return torch.empty(0, dtype=torch.long, device=weights.device)
    u = torch.rand(w.shape, dtype=torch.float64, device=w.device, generator=generator).clamp_min(1e-300)
    keys = torch.where(w > 0, torch.log(u) / w, torch.full_like(w, -float("inf")))  # log(u^(1/w))
    return torch.topk(keys, n).indices


def quaternion_vec_multiply(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """quat_vec.rs: rotate v [n,3] by (not necessarily unit) q [n,4] = (w,x,y,z)."""
    qw, qx, qy, qz = q[:, 0:1], q[:, 1:2], q[:, 2:3], q[:, 3:4]
    vx, vy, vz = v[:, 0:1], v[:, 1:2], v[:, 2:3]
    qw2, qx2, qy2, qz2 = qw * qw, qx * qx, qy * qy, qz * qz
    xy, xz, yz, wx, wy, wz = qx * qy, qx * qz, qy * qz, qw * qx, qw * qy, qw * qz
    x = (qw2 + qx2 - qy2 - qz2) * vx + (xy * vy + xz * vz + wy * vz - wz * vy) * 2.0
    y = (qw2 - qx2 + qy2 - qz2) * vy + (xy * vx + yz * vz + wz * vx - wx * vz) * 2.0
    z = (qw2 - qx2 - qy2 + qz2) * vz + (xz * vx + yz * vy + wx * vy - wy * vx) * 2.0
    return torch.cat([x, y, z], 1)


def bounds_from_pos_torch(percentile: float, means: torch.Tensor) -> BoundingBox:
    """splat_init.rs:130-160 on the device (one sort per axis, 6 scalars read back)."""
    vals = []
    for a in range(3):
        v = means[:, a]
        v = torch.sort(v[torch.isfinite(v)]).values
        n = v.numel()
        if n == 0:
            return BoundingBox(np.zeros(3, np.float32), np.ones(3, np.float32))
        lo = int((np.float32(1.0) - np.float32(percentile)) / np.float32(2.0) * np.float32(n))   # all in f32, like the reference
        hi = min(n - 1, int((np.float32(1.0) + np.float32(percentile)) / np.float32(2.0) * np.float32(n)))
        vals.append(torch.stack([v[lo], v[hi]]))
    mm = torch.stack(vals).cpu().numpy().astype(np.float32)  # [3,2]
    return BoundingBox((mm[:, 1] + mm[:, 0]) / 2.0, (mm[:, 1] - mm[:, 0]) / 2.0)



def refine_reference(trainer, iteration: int, splats) -> RefineStats:
    """The torch version of refine(): `trainer` supplies config, bounds, _state and a seeded torch.Generator (_gen)."""
    self = trainer
    cfg = self.config
    if self._state is None:
        raise RuntimeError("Can only refine after optimizer is initialized")
    st = self._state
    dev = splats.transforms.device
    max_allowed = float(np.max(self.bounds.extent)) * 100.0

    # ---- prune mask (train.rs:487-535)
    opac = torch.sigmoid(splats.raw_opacities)
    alpha_mask = opac < MIN_OPACITY
    scale_big = (splats.transforms[:, 7:10].exp() > max_allowed).any(1)
    center = torch.tensor(self.bounds.center, dtype=torch.float32, device=dev).reshape(1, 3)
    bound_mask = ((splats.transforms[:, 0:3] - center).abs() > max_allowed).any(1)
    non_finite = (~torch.isfinite(splats.transforms)).any(1) | (~torch.isfinite(splats.sh_coeffs.flatten(1))).any(1) \
        | ~torch.isfinite(splats.raw_opacities)
    num_non_finite = int(non_finite.sum().item())
    prune = alpha_mask | scale_big | bound_mask | non_finite

    # ---- prune_points (train.rs:848-893)
    keep = (~prune).nonzero(as_tuple=False).squeeze(1)
    n0 = splats.num_splats()
    pruned = 0
    if 0 < keep.numel() < n0:
        pruned = n0 - keep.numel()
        splats.transforms = splats.transforms.index_select(0, keep)
        splats.sh_coeffs = splats.sh_coeffs.index_select(0, keep)
        splats.raw_opacities = splats.raw_opacities.index_select(0, keep)
        for k in ("m_t", "v_t", "m_sh", "v_sh", "m_o", "v_o", "refine_norm", "vis_weight", "max_screen"):
            st[k] = st[k].index_select(0, keep)
    n = splats.num_splats()
    vis_mask = st["vis_weight"] > 0

    split = torch.zeros(n, dtype=torch.bool, device=dev)
    # ---- replace dead gaussians, weighted by opacity x visibility (train.rs:544-556)
    if pruned > 0:
        wts = torch.sigmoid(splats.raw_opacities) * vis_mask.float()
        split[multinomial_sample(wts, pruned, self._gen)] = True
    # ---- force-split splats that are too big on screen (train.rs:562-586), capped by max_splats
    pre = int(split.sum().item())
    if cfg.split_at_screen_size > 0.0:
        cand = ((st["max_screen"] > cfg.split_at_screen_size) & vis_mask & ~split).nonzero(as_tuple=False).squeeze(1)
        budget = max(0, cfg.max_splats - (n + pre))
        split[cand[:budget]] = True
    num_oversized = int(split.sum().item()) - pre
    # ---- growth: sample among splats whose refine weight is above the threshold (train.rs:590-632)
    pre_grad = int(split.sum().item())
    if iteration < cfg.growth_stop_iter:                     # train.rs:591
        above = (st["refine_norm"] > cfg.growth_grad_threshold) & vis_mask
        threshold_count = int(above.sum().item())
        grow = max(0, int(np.floor(np.float32(threshold_count) * np.float32(cfg.growth_select_fraction) + np.float32(0.5))) - pruned)   # f32::round
        grow = min(grow, max(0, cfg.max_splats - (n + pre_grad)))
        if grow > 0:
            # sampled independently of earlier picks, like the reference's HashSet union
            split[multinomial_sample(above.float() * st["refine_norm"], grow, self._gen)] = True
    num_high_grad = int(split.sum().item()) - pre_grad
    inds = split.nonzero(as_tuple=False).squeeze(1)
    refine_count = inds.numel()

    # ---- refine_splats (train.rs:665-821)
    if refine_count > 0:
        cur = splats.transforms.index_select(0, inds)
        cur_means, rots_raw, cur_log_scale = cur[:, 0:3], cur[:, 3:7], cur[:, 7:10]
        cur_rots = rots_raw / rots_raw.pow(2).sum(1, keepdim=True).sqrt().clamp_min(1e-32)
        cur_sh = splats.sh_coeffs.index_select(0, inds)
        cur_raw_opac = splats.raw_opacities.index_select(0, inds)
        cur_scales = cur_log_scale.exp()
        inv_opac = 1.0 - torch.sigmoid(cur_raw_opac)
        new_opac = (1.0 - inv_opac.pow(FRAC_1_SQRT_2)).clamp(MIN_OPACITY, 1.0 - MIN_OPACITY)
        new_raw_opac = torch.log(new_opac / (1.0 - new_opac))
        sq = cur_scales.pow(2)
        ratio = sq / sq.max(1, keepdim=True).values.clamp_min(1e-30)
        if cfg.split_at_screen_size > 0.0:
            k_max = (st["max_screen"].index_select(0, inds).unsqueeze(1).clamp_min(1e-6).reciprocal()
                     * cfg.split_at_screen_size).clamp_max(FRAC_1_SQRT_2)
            k_axis = -(ratio * (1.0 - k_max)) + 1.0
        else:
            k_axis = -(ratio * (1.0 - FRAC_1_SQRT_2)) + 1.0
        offset_local = (1.0 - k_axis.pow(2)).clamp_min(0.0).sqrt() * cur_scales
        samples = quaternion_vec_multiply(cur_rots, offset_local)
        new_log_scales = cur_log_scale + k_axis.log()
        # parents move to mean - offset and shrink; children sit at mean + offset
        splats.transforms[inds, 0:3] = cur_means - samples
        splats.transforms[inds, 7:10] = new_log_scales
        splats.raw_opacities[inds] = new_raw_opac
        children = torch.cat([cur_means + samples, cur_rots, new_log_scales], 1)
        splats.transforms = torch.cat([splats.transforms, children], 0)
        splats.sh_coeffs = torch.cat([splats.sh_coeffs, cur_sh], 0)
        splats.raw_opacities = torch.cat([splats.raw_opacities, new_raw_opac], 0)
        # both halves of a split restart with zero Adam moments
        for k in ("m_t", "v_t", "m_sh", "v_sh", "m_o", "v_o"):
            st[k][inds] = 0
            st[k] = torch.cat([st[k], torch.zeros((refine_count,) + tuple(st[k].shape[1:]), dtype=torch.float32, device=dev)], 0)

    # ---- opacity decay (train.rs:808-816)
    train_t = min(max(iteration / float(cfg.total_train_iters), 0.0), 1.0)
    minus_opac = cfg.opac_decay * (1.0 - train_t)
    o = (torch.sigmoid(splats.raw_opacities) - minus_opac).clamp(1e-12, 1.0 - 1e-12)
    splats.raw_opacities = torch.log(o / (1.0 - o))

    # ---- bounds, refine record restart (train.rs:637-640, 442-445)
    self.bounds = bounds_from_pos_torch(BOUND_PERCENTILE, splats.transforms[:, 0:3])
    n_new = splats.num_splats()
    for k in ("refine_norm", "vis_weight", "max_screen"):
        st[k] = torch.zeros(n_new, dtype=torch.float32, device=dev)
    splats.transforms = splats.transforms.contiguous()
    splats.sh_coeffs = splats.sh_coeffs.contiguous()
    splats.raw_opacities = splats.raw_opacities.contiguous()
    return RefineStats(num_added=refine_count, num_split_oversized=num_oversized, num_split_high_grad=num_high_grad,
                       num_pruned=pruned, num_pruned_non_finite=num_non_finite, total_splats=n_new)
