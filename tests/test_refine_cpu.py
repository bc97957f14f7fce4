"""refine() semantics (brush-train/src/train.rs:431-893, multinomial.rs, quat_vec.rs) on CPU tensors, through the test-side
torch restatement tests/refine_ref.py (the product path is bg_refine, compared against it in tests/test_gpu_refine.py).  The reference pins none of
this numerically (unseeded RNG); the tests check its own unit tests' properties (multinomial.rs tests,
quat_vec.rs tests) and the invariants stated in the code/comments."""
import math

import numpy as np
import pytest
import torch

from brush_b200.train import BoundingBox, RefineStats, SplatTrainer, Splats, TrainConfig, bounds_from_pos
from refine_ref import bounds_from_pos_torch as bounds_from_pos_device
from refine_ref import multinomial_sample, quaternion_vec_multiply, refine_reference


def test_multinomial_sampling_properties():
    """multinomial.rs:28-85."""
    g = torch.Generator().manual_seed(0)
    w = torch.tensor([0.1, 0.3, 0.4, 0.2])
    s = multinomial_sample(w, 3, g)
    assert s.numel() == 3 and len(set(s.tolist())) == 3 and all(0 <= i < 4 for i in s.tolist())
    assert multinomial_sample(torch.tensor([1.0]), 1, g).tolist() == [0]
    s = multinomial_sample(torch.tensor([0.5, float("nan"), 0.3, 0.2]), 2, g)
    assert s.numel() == 2 and 1 not in s.tolist()
    assert multinomial_sample(torch.zeros(3), 1, g).numel() == 0
    # proportional to the weights (without replacement, n=1)
    w = torch.tensor([1.0, 3.0, 6.0])
    hits = torch.zeros(3)
    for _ in range(4000):
        hits[multinomial_sample(w, 1, g)] += 1
    assert torch.allclose(hits / 4000, w / w.sum(), atol=0.03)


def test_quaternion_vec_multiply():
    q = torch.tensor([[1.0, 0, 0, 0], [math.cos(math.pi / 4), 0, 0, math.sin(math.pi / 4)]])
    v = torch.tensor([[1.0, 2, 3], [1.0, 0, 0]])
    r = quaternion_vec_multiply(q, v)
    assert torch.allclose(r[0], v[0]) and torch.allclose(r[1], torch.tensor([0.0, 1.0, 0.0]), atol=1e-6)


def _trainer(n, k=4, seed=0, **cfg):
    g = torch.Generator().manual_seed(seed)
    means = torch.rand(n, 3, generator=g) * 4 - 2
    quats = torch.randn(n, 4, generator=g)
    ls = torch.log(torch.rand(n, 3, generator=g) * 0.05 + 0.01)
    splats = Splats(torch.cat([means, quats, ls], 1).contiguous(), torch.randn(n, k, 3, generator=g) * 0.2,
                    torch.rand(n, generator=g) * 4 - 1)
    bounds = bounds_from_pos(0.8, means.numpy())
    tr = SplatTrainer(TrainConfig(**cfg), None, bounds)
    tr._ensure_state(splats)
    st = tr._state
    for key in ("m_t", "v_t", "m_sh", "v_sh", "m_o", "v_o"):
        st[key] += 1.0  # non-zero moments so that resets are observable
    st["vis_weight"] += 1.0
    tr.step_count = 200
    return tr, splats


def test_refine_prunes_dead_and_nonfinite_and_reuses_the_budget():
    tr, sp = _trainer(1000)
    sp.raw_opacities[:50] = -8.0            # sigmoid < 1/255
    sp.transforms[50:55, 0] = float("nan")
    sp.transforms[55:60, 7] = 20.0          # scale far above 100 * extent
    n0 = sp.num_splats()
    stats = refine_reference(tr, 200, sp)
    assert isinstance(stats, RefineStats)
    assert stats.num_pruned == 60 and stats.num_pruned_non_finite == 5
    assert stats.num_added >= 60            # pruned budget is re-used by splitting survivors
    assert stats.total_splats == n0 - 60 + stats.num_added == sp.num_splats()
    st = tr._state
    for key in ("m_t", "v_t", "m_sh", "v_sh", "m_o", "v_o", "refine_norm", "vis_weight", "max_screen"):
        assert st[key].shape[0] == sp.num_splats()
    assert torch.isfinite(sp.transforms).all() and torch.isfinite(sp.raw_opacities).all()
    assert (st["refine_norm"] == 0).all() and (st["vis_weight"] == 0).all()   # record restarts
    assert sp.transforms.is_contiguous() and sp.sh_coeffs.shape[0] == sp.num_splats()


def test_split_geometry_opacity_and_moments():
    tr, sp = _trainer(200, opac_decay=0.0)
    tr._state["refine_norm"][:] = 0.0
    tr._state["refine_norm"][10] = 1.0      # one splat above the growth threshold
    tr._state["max_screen"][:] = 0.01
    before = Splats(sp.transforms.clone(), sp.sh_coeffs.clone(), sp.raw_opacities.clone())
    stats = refine_reference(tr, 100, sp)
    # growth = round(0.25 * 1) - 0 = 0 -> nothing split; raise the fraction to force the split
    assert stats.num_added == 0
    tr, sp = _trainer(200, opac_decay=0.0, growth_select_fraction=1.0)
    tr._state["refine_norm"][:] = 0.0
    tr._state["refine_norm"][10] = 1.0
    tr._state["max_screen"][:] = 0.01
    before = Splats(sp.transforms.clone(), sp.sh_coeffs.clone(), sp.raw_opacities.clone())
    stats = refine_reference(tr, 100, sp)
    assert stats.num_added == 1 and stats.num_split_high_grad == 1 and sp.num_splats() == 201
    parent, child, old = sp.transforms[10], sp.transforms[200], before.transforms[10]
    # centroid preserved, children symmetric about the old mean
    assert torch.allclose((parent[:3] + child[:3]) / 2, old[:3], atol=1e-6)
    # max axis shrinks by 1/sqrt(2), the others less (k_axis = 1 - ratio (1 - k))
    shrink = (parent[7:10] - old[7:10]).exp()
    assert torch.allclose(shrink.min(), torch.tensor(1 / math.sqrt(2)), atol=1e-5) and (shrink <= 1 + 1e-6).all()
    assert torch.allclose(parent[7:10], child[7:10])
    # child opacity 1 - (1 - o)^(1/sqrt 2), both halves
    o = torch.sigmoid(before.raw_opacities[10])
    exp_o = (1 - (1 - o) ** (1 / math.sqrt(2))).clamp(1 / 255, 1 - 1 / 255)
    assert torch.allclose(torch.sigmoid(sp.raw_opacities[10]), exp_o, atol=1e-6)
    assert torch.allclose(sp.raw_opacities[200], sp.raw_opacities[10])
    assert torch.equal(sp.sh_coeffs[200], before.sh_coeffs[10])
    # rotation of the child is the normalised parent rotation
    assert torch.allclose(child[3:7], old[3:7] / old[3:7].norm(), atol=1e-6)
    st = tr._state
    for key in ("m_t", "v_t", "m_sh", "v_sh", "m_o", "v_o"):
        assert float(st[key][10].abs().sum()) == 0 and float(st[key][200].abs().sum()) == 0
        assert float(st[key][11].abs().sum()) > 0     # untouched splats keep their moments


def test_oversized_force_split_and_growth_stop():
    tr, sp = _trainer(300, growth_stop_iter=100)
    tr._state["max_screen"][:5] = 0.9
    tr._state["refine_norm"][:] = 1.0        # would all grow, but iteration >= growth_stop_iter
    stats = refine_reference(tr, 150, sp)
    assert stats.num_split_oversized == 5 and stats.num_split_high_grad == 0 and stats.num_added == 5
    # max_splats caps the force split
    tr, sp = _trainer(300, max_splats=302)
    tr._state["max_screen"][:5] = 0.9
    tr._state["refine_norm"][:] = 0.0
    stats = refine_reference(tr, 10, sp)
    assert stats.num_split_oversized == 2 and sp.num_splats() == 302


def test_opacity_decay_and_bounds_update():
    tr, sp = _trainer(500, total_train_iters=1000)
    tr._state["refine_norm"][:] = 0.0
    tr._state["max_screen"][:] = 0.0
    o0 = torch.sigmoid(sp.raw_opacities).clone()
    keep = o0 >= 1 / 255
    refine_reference(tr, 250, sp)
    o1 = torch.sigmoid(sp.raw_opacities)
    assert torch.allclose(o1, (o0[keep] - 0.004 * 0.75).clamp(1e-12, 1 - 1e-12), atol=1e-6)
    b = bounds_from_pos(0.8, sp.transforms[:, :3].numpy())
    np.testing.assert_allclose(tr.bounds.extent, b.extent, rtol=1e-6)
    d = bounds_from_pos_device(0.8, sp.transforms[:, :3])
    np.testing.assert_allclose(d.center, b.center, rtol=1e-6, atol=1e-7)


def test_bounds_from_pos_degenerate_inputs():
    """splat_init.rs:245-295: all-NaN, empty, mixed and one-axis-NaN inputs give finite boxes (unit-box fallback)."""
    import numpy as np
    for fn in (bounds_from_pos, lambda p, m: bounds_from_pos_device(p, torch.from_numpy(m))):
        for means in (np.full((10, 3), np.nan, np.float32), np.zeros((0, 3), np.float32)):
            bb = fn(0.8, means)
            assert np.isfinite(bb.center).all() and np.isfinite(bb.extent).all()
            assert (bb.center == 0).all() and (bb.extent == 1).all()
        mixed = np.full((100, 3), np.nan, np.float32)
        mixed[1::2] = np.arange(1, 100, 2, dtype=np.float32)[:, None]
        bb = fn(0.8, mixed)
        assert np.isfinite(bb.center).all() and 0.0 < bb.extent[0] < 100.0
        one_axis = np.stack([np.arange(50), np.full(50, np.nan), np.arange(50)], 1).astype(np.float32)
        bb = fn(0.8, one_axis)
        assert np.isfinite(bb.center).all() and np.isfinite(bb.extent).all()


def test_restatement_philox_matches_the_published_known_answers():
    """The numpy Philox4x32-10 that tests/test_gpu_refine.py uses to re-draw the device's uniforms, against the Random123
    known-answer vectors (kat_vectors: philox4x32 10 rounds) -- so the restatement itself is pinned."""
    import numpy as np
    import test_gpu_refine as tgr
    z = np.zeros(1, np.uint64)
    assert [int(x[0]) for x in tgr.philox4x32_10(z, z, z, z, 0, 0)] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = np.full(1, 0xFFFFFFFF, np.uint64)
    assert [int(x[0]) for x in tgr.philox4x32_10(f, f, f, f, 0xFFFFFFFF, 0xFFFFFFFF)] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    c = [np.array([v], np.uint64) for v in (0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344)]
    assert [int(x[0]) for x in tgr.philox4x32_10(*c, 0xA4093822, 0x299F31D0)] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    u = tgr.uniform01(9, 800, np.arange(4096, dtype=np.uint32))
    assert u.dtype == np.float32 and 0.0 < u.min() and u.max() < 1.0 and abs(float(u.mean()) - 0.5) < 0.02
