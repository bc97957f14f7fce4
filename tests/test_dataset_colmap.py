"""COLMAP text loader and packed-view producer, after the reference's own tests
(brush-dataset/src/formats/colmap.rs:392-590, scene.rs:164+)."""
import math
import os

import numpy as np
import pytest

from brush_b200 import camera as cm
from brush_b200 import dataset as ds

IMG_W, IMG_H, FX, FY, CX, CY = 64, 48, 80.0, 70.0, 30.0, 20.0


def _img1_w2c():
    a = math.pi / 4.0                       # Quat::from_rotation_y(pi/2) = (w, x, y, z) = (cos a, 0, sin a, 0)
    return (math.cos(a), 0.0, math.sin(a), 0.0), (1.0, 0.0, 2.0)


def _write_dataset(root):
    from PIL import Image
    sparse = os.path.join(root, "sparse", "0")
    os.makedirs(sparse)
    open(os.path.join(sparse, "cameras.txt"), "w").write(
        "# Camera list with one line of data per camera:\n#   CAMERA_ID, MODEL, WIDTH, HEIGHT, PARAMS[]\n"
        f"1 PINHOLE {IMG_W} {IMG_H} {FX} {FY} {CX} {CY}\n")
    q, t = _img1_w2c()
    open(os.path.join(sparse, "images.txt"), "w").write(
        "# Image list with two lines of data per image:\n#   IMAGE_ID, QW, QX, QY, QZ, TX, TY, TZ, CAMERA_ID, NAME\n"
        "#   POINTS2D[] as (X, Y, POINT3D_ID)\n"
        "1 1.0 0.0 0.0 0.0 1.0 2.0 3.0 1 img0.png\n\n"
        f"2 {q[0]} {q[1]} {q[2]} {q[3]} {t[0]} {t[1]} {t[2]} 1 img1.png\n10.0 20.0 1 30.0 40.0 -1\n"
        "3 1.0 0.0 0.0 0.0 0.0 0.0 0.0 1 img2.png\n\n"
        "4 1.0 0.0 0.0 0.0 5.0 5.0 5.0 1 missing.png\n\n")
    open(os.path.join(sparse, "points3D.txt"), "w").write(
        "# 3D point list with one line of data per point:\n"
        "1 1.5 2.5 3.5 255 0 0 0.5 1 0\n2 -1.0 0.0 1.0 0 255 0 0.5 2 1\n3 0.0 1.0 0.0 0 0 255 0.5 3 0\n"
        "4 2.0 2.0 2.0 128 128 128 0.5 1 1\n")
    os.makedirs(os.path.join(root, "images"))
    for name in ("img0.png", "img1.png", "img2.png"):
        Image.fromarray(np.full((3, 4, 3), (10, 20, 30), np.uint8)).save(os.path.join(root, "images", name))


def test_loads_text_model(tmp_path):
    _write_dataset(str(tmp_path))
    r = ds.load_colmap_text(str(tmp_path))
    assert len(r.train) == 3 and r.eval == [] and len(r.warnings) == 1 and "missing.png" in r.warnings[0]
    cam = r.train[0].camera
    assert cam.camera_model == cm.PINHOLE
    assert abs(cam.fov_x - cm.focal_to_fov(FX, IMG_W)) < 1e-6 and abs(cam.fov_y - cm.focal_to_fov(FY, IMG_H)) < 1e-6
    fx, fy = cam.focal(IMG_W, IMG_H)
    cx, cy = cam.center(IMG_W, IMG_H)
    assert abs(fx - FX) < 1e-3 and abs(fy - FY) < 1e-3 and abs(cx - CX) < 1e-3 and abs(cy - CY) < 1e-3
    # sorted by name; img0: identity rotation, w2c translation (1,2,3) -> camera at (-1,-2,-3)
    np.testing.assert_allclose(r.train[0].camera.position, (-1, -2, -3), atol=1e-4)
    np.testing.assert_allclose(np.abs(r.train[0].camera.rotation), (0, 0, 0, 1), atol=1e-6)
    np.testing.assert_allclose(r.train[2].camera.position, (0, 0, 0), atol=1e-6)
    # img1: -R^T (1, 0, 2) = (2, 0, -1); rotation is the inverse of the written one
    np.testing.assert_allclose(r.train[1].camera.position, (2.0, 0.0, -1.0), atol=1e-4)
    q, _ = _img1_w2c()
    np.testing.assert_allclose(r.train[1].camera.rotation, (-q[1], -q[2], -q[3], q[0]), atol=1e-6)
    # the camera built from it maps the camera centre to the origin of its local frame
    vm = r.train[1].camera.world_to_local().reshape(4, 3)
    np.testing.assert_allclose(vm[:3].T @ np.array(r.train[1].camera.position) + vm[3], 0, atol=1e-5)
    init = r.init_splat
    assert init.num_splats() == 4 and (init.means[0] == [1.5, 2.5, 3.5]).all()
    want = (np.array([1, 0, 0], np.float32) - np.float32(0.5)) / np.float32(0.2820947917738781)
    np.testing.assert_array_equal(init.sh_coeffs[0, 0], want)
    packed, has_alpha = r.train[0].load_packed()
    assert packed.shape == (3, 4) and not has_alpha and (packed.view(np.uint32) == (10 | 20 << 8 | 30 << 16 | 255 << 24)).all()


def test_splits_eval_views(tmp_path):
    _write_dataset(str(tmp_path))
    r = ds.load_colmap_text(str(tmp_path), eval_split_every=2)
    assert len(r.train) == 1 and len(r.eval) == 2
    np.testing.assert_allclose(r.train[0].camera.position, (2.0, 0.0, -1.0), atol=1e-4)
    np.testing.assert_allclose(r.eval[0].camera.position, (-1, -2, -3), atol=1e-4)
    np.testing.assert_allclose(r.eval[1].camera.position, (0, 0, 0), atol=1e-6)


def test_colmap_camera_models_map_like_the_reference():
    c = ds.ColmapCamera(1, "OPENCV", 100, 80, [90, 91, 50, 40, -0.1, 0.02, 1e-3, -2e-3])
    assert ds.build_camera_model(c) == (cm.RADIAL_TANGENTIAL_8, pytest.approx((-0.1, 0.02, 0, 0, 0, 0, 1e-3, -2e-3)))
    c = ds.ColmapCamera(1, "FULL_OPENCV", 100, 80, [90, 91, 50, 40, 1, 2, 3, 4, 5, 6, 7, 8])
    assert ds.build_camera_model(c) == (cm.RADIAL_TANGENTIAL_8, (1, 2, 5, 6, 7, 8, 3, 4))
    c = ds.ColmapCamera(1, "OPENCV_FISHEYE", 100, 80, [90, 91, 50, 40, 1, 2, 3, 4])
    assert ds.build_camera_model(c) == (cm.KANNALA_BRANDT_4, (1, 2, 3, 4))
    c = ds.ColmapCamera(1, "THIN_PRISM_FISHEYE", 100, 80, [90, 91, 50, 40, 1, 2, 3, 4, 5, 6, 7, 8])
    assert ds.build_camera_model(c) == (cm.THIN_PRISM_FISHEYE, (1, 2, 5, 6, 3, 4, 7, 8))
    c = ds.ColmapCamera(1, "SIMPLE_RADIAL", 100, 80, [90, 50, 40, 0.05])
    assert c.focal() == (90, 90) and c.principal_point() == (50, 40)
    assert ds.build_camera_model(c)[0] == cm.RADIAL_TANGENTIAL_8
    assert ds.build_camera_model(ds.ColmapCamera(1, "FOV", 1, 1, [1, 1, 0, 0, 0.5])) == (cm.PINHOLE, ())
    cam = ds.camera_from_colmap(ds.ColmapCamera(1, "OPENCV_FISHEYE", 640, 480, [300, 300, 320, 240, -0.01, 0.003, 0, 0]),
                                ds.ColmapImage(1, (1, 0, 0, 0), (0, 0, 0), 1, "a.png"))
    f = cm.fov_to_focal(cam.fov_x, 640, cam.camera_model, cam.model_params)
    assert abs(f - 300) < 1e-6


def test_view_to_packed_data_and_premultiplication():
    rgb = np.arange(2 * 3 * 3, dtype=np.uint8).reshape(2, 3, 3)
    p, has_alpha = ds.view_to_packed_data(rgb)
    assert not has_alpha and p.dtype == np.int32
    u = p.view(np.uint32)
    assert (u & 0xFF == rgb[..., 0]).all() and ((u >> 8) & 0xFF == rgb[..., 1]).all() and (u >> 24 == 255).all()
    rgba = np.array([[[200, 100, 50, 128], [255, 255, 255, 0], [7, 8, 9, 255]]], np.uint8)
    masked, ha = ds.view_to_packed_data(rgba, ds.ALPHA_MASKED)
    assert ha and (masked.view(np.uint32)[0] == [200 | 100 << 8 | 50 << 16 | 128 << 24, 0x00FFFFFF, 7 | 8 << 8 | 9 << 16 | 255 << 24]).all()
    pre, _ = ds.view_to_packed_data(rgba, ds.ALPHA_TRANSPARENT)
    mul = lambda c, a: (c * a + 127) // 255
    assert pre.view(np.uint32)[0, 0] == (mul(200, 128) | mul(100, 128) << 8 | mul(50, 128) << 16 | 128 << 24)
    assert pre.view(np.uint32)[0, 1] == 0 and pre.view(np.uint32)[0, 2] == masked.view(np.uint32)[0, 2]
    with pytest.raises(ValueError):
        ds.view_to_packed_data(np.zeros((2, 2), np.uint8))


def test_scene_loader_cycles_through_views(tmp_path):
    _write_dataset(str(tmp_path))
    r = ds.load_colmap_text(str(tmp_path))
    loader = ds.SceneLoader(r.train, seed=1, threads=1)     # one loader task: every view once per epoch
    seen = {loader.next_batch().camera.position for _ in range(3)}
    assert len(seen) == 3
    b = loader.next_batch()
    assert tuple(b.img_packed.shape) == (3, 4) and not b.has_alpha
    loader.close()


def test_scene_loader_threads_and_cache(tmp_path):
    """scene_loader.rs:13-58,67-110: several loader tasks feed one bounded queue; a packed batch that fits the cache
    budget is handed out again as the SAME buffer (never rewritten), one that does not fit is re-made every time."""
    _write_dataset(str(tmp_path))
    r = ds.load_colmap_text(str(tmp_path))
    loader = ds.SceneLoader(r.train, seed=3, threads=3)
    got = [loader.next_batch() for _ in range(40)]
    assert {b.camera.position for b in got} == {v.camera.position for v in r.train}
    by_view = {}
    for b in got:
        by_view.setdefault(b.camera.position, []).append(b.img_packed)
    for tensors in by_view.values():
        assert len({t.data_ptr() for t in tensors}) == 1          # cached: one buffer per view
    loader.close()
    tiny = ds.SceneLoader(r.train, seed=3, threads=2, cache_bytes=0)   # nothing fits: fresh buffers, same contents
    a = [tiny.next_batch() for _ in range(12)]
    ref = {v.camera.position: v.load_packed()[0] for v in r.train}
    for b in a:
        np.testing.assert_array_equal(b.img_packed.numpy().view(np.uint32), ref[b.camera.position].view(np.uint32))
    tiny.close()


def test_binary_model_equals_text_model(tmp_path):
    """colmap-reader's binary layout (lib.rs:278-300, 389-470, 546-600): same dataset, written as .bin."""
    import struct
    txt_root, bin_root = tmp_path / "t", tmp_path / "b"
    txt_root.mkdir(); bin_root.mkdir()
    _write_dataset(str(txt_root))
    _write_dataset(str(bin_root))
    sparse = bin_root / "sparse" / "0"
    for f in ("cameras.txt", "images.txt", "points3D.txt"):
        os.remove(sparse / f)
    (sparse / "cameras.bin").write_bytes(struct.pack("<QiiQQ4d", 1, 1, 1, IMG_W, IMG_H, FX, FY, CX, CY))
    q, t = _img1_w2c()
    imgs = [(1, (1.0, 0, 0, 0), (1.0, 2.0, 3.0), "img0.png", 0), (2, q, t, "img1.png", 2), (3, (1.0, 0, 0, 0), (0, 0, 0), "img2.png", 0),
            (4, (1.0, 0, 0, 0), (5.0, 5.0, 5.0), "missing.png", 0)]
    blob = struct.pack("<Q", len(imgs))
    for iid, qq, tt, name, npts in imgs:
        blob += struct.pack("<i7di", iid, *qq, *tt, 1) + name.encode() + b"\0" + struct.pack("<Q", npts)
        blob += b"".join(struct.pack("<ddq", 10.0 * j, 20.0, -1) for j in range(npts))
    (sparse / "images.bin").write_bytes(blob)
    pts = [(1, (1.5, 2.5, 3.5), (255, 0, 0), 1), (2, (-1.0, 0.0, 1.0), (0, 255, 0), 2), (3, (0.0, 1.0, 0.0), (0, 0, 255), 0),
           (4, (2.0, 2.0, 2.0), (128, 128, 128), 1)]
    blob = struct.pack("<Q", len(pts))
    for pid, xyz, rgb, track in pts:
        blob += struct.pack("<q3d3BdQ", pid, *xyz, *rgb, 0.5, track) + b"".join(struct.pack("<ii", 1, j) for j in range(track))
    (sparse / "points3D.bin").write_bytes(blob)
    a, b = ds.load_colmap(str(txt_root)), ds.load_colmap(str(bin_root))
    assert len(a.train) == len(b.train) == 3 and len(b.warnings) == 1
    for va, vb in zip(a.train, b.train):
        np.testing.assert_allclose(va.camera.position, vb.camera.position, atol=1e-6)
        np.testing.assert_allclose(va.camera.rotation, vb.camera.rotation, atol=1e-6)
        assert abs(va.camera.fov_x - vb.camera.fov_x) < 1e-12 and va.camera.center_uv == vb.camera.center_uv
    np.testing.assert_array_equal(a.init_splat.means, b.init_splat.means)
    np.testing.assert_array_equal(a.init_splat.sh_coeffs, b.init_splat.sh_coeffs)
    with pytest.raises(ValueError):
        ds.read_cameras_binary(struct.pack("<QiiQQ", 1, 1, 99, 4, 4))
    with pytest.raises(ValueError):
        ds.read_images_binary(struct.pack("<Qi7di", 1, 1, 1, 0, 0, 0, 0, 0, 0, 1) + b"unterminated")


# ---------------------------------------------------------------------------------------------- masks
def test_find_mask_path_rules():
    """brush-dataset/src/formats/mod.rs:197-272 restated."""
    f = ds.find_mask_path
    assert f(["images/img.png", "masks/img.png"], "images/img.png") == "masks/img.png"
    assert f(["images/img.jpeg", "masks/img.png"], "images/img.jpeg") == "masks/img.png"          # other extension
    assert f(["images/foo.png", "masks/foo.png.mask"], "images/foo.png") == "masks/foo.png.mask"  # img.png.*
    assert f(["images/bar.jpeg", "masks/bar.mask.png"], "images/bar.jpeg") == "masks/bar.mask.png"  # img.mask.*
    assert f(["images/foo/bar/img.png", "masks/foo/bar/img.png"], "images/foo/bar/img.png") == "masks/foo/bar/img.png"
    assert f(["images/baz/img.png", "masks/foo/img.png"], "images/baz/img.png") is None           # wrong sub-directory
    assert f(["images/IMG.PNG", "masks/img.png"], "images/IMG.PNG") == "masks/img.png"            # case-insensitive
    assert f(["images/img.png", "other/img.png"], "images/img.png") is None                       # not under masks/
    assert f(["images/img.png", "MASKS/img.png"], "images/img.png") == "MASKS/img.png"


def _mask_view(tmp_path, mask_img, invert):
    from PIL import Image
    Image.new("RGB", (4, 2), (10, 20, 30)).save(tmp_path / "img.png")
    mask_img.save(tmp_path / "mask.png")
    v = ds.SceneView(cm.Camera(), str(tmp_path / "img.png"), str(tmp_path / "mask.png"), invert)
    packed, has_alpha = v.load_packed(v.default_alpha_mode())
    assert has_alpha and packed.shape == (2, 4)
    return packed.view(np.uint32)


def test_mask_becomes_alpha(tmp_path):
    """load_image.rs:256-262: the grey mask is the alpha channel, colours stay intact (Masked: no premultiplication)."""
    from PIL import Image
    vals = np.array([i * 30 for i in range(8)], np.uint8).reshape(2, 4)
    px = _mask_view(tmp_path, Image.fromarray(vals, "L"), False)
    assert np.array_equal((px >> 24).astype(np.uint8), vals)
    assert np.array_equal(px & 0xFFFFFF, np.full((2, 4), 10 | (20 << 8) | (30 << 16), np.uint32))


def test_inverted_mask_flips_alpha(tmp_path):
    """load_image.rs:264-270."""
    from PIL import Image
    vals = np.array([i * 30 for i in range(8)], np.uint8).reshape(2, 4)
    px = _mask_view(tmp_path, Image.fromarray(vals, "L"), True)
    assert np.array_equal((px >> 24).astype(np.uint8), 255 - vals)


def test_mask_with_alpha_channel_and_other_size(tmp_path):
    """load_image.rs:85-102: a mask that has an alpha channel contributes THAT channel; a mask of another size is
    resized to the image."""
    from PIL import Image
    rgba = np.zeros((2, 4, 4), np.uint8)
    rgba[..., 0] = 255                       # the colour channels must be ignored
    rgba[..., 3] = np.array([[0, 50, 100, 150], [200, 250, 5, 15]], np.uint8)
    px = _mask_view(tmp_path, Image.fromarray(rgba, "RGBA"), False)
    assert np.array_equal((px >> 24).astype(np.uint8), rgba[..., 3])
    big = np.repeat(np.repeat(np.array([[0, 255], [255, 0]], np.uint8), 4, axis=0), 8, axis=1)    # 8 x 16 for a 2 x 4 image
    px = _mask_view(tmp_path, Image.fromarray(big, "L"), False)
    a = (px >> 24).astype(np.int32)
    assert a.shape == (2, 4) and a[0, 0] < 64 and a[0, 3] > 192 and a[1, 0] > 192 and a[1, 3] < 64


def test_colmap_loader_attaches_masks(tmp_path):
    from PIL import Image
    root = str(tmp_path)
    _write_dataset(root)
    os.makedirs(os.path.join(root, "masks"))
    Image.fromarray(np.full((IMG_H, IMG_W), 200, np.uint8), "L").save(os.path.join(root, "masks", "img1.png"))
    r = ds.load_colmap(root)
    by_name = {os.path.basename(v.image_path): v for v in r.train + r.eval}
    assert by_name["img1.png"].mask_path == os.path.join(root, "masks", "img1.png") and not by_name["img1.png"].invert_mask
    assert by_name["img0.png"].mask_path is None and by_name["img0.png"].default_alpha_mode() == ds.ALPHA_TRANSPARENT
    assert by_name["img1.png"].default_alpha_mode() == ds.ALPHA_MASKED
    packed, has_alpha = by_name["img1.png"].load_packed(ds.ALPHA_MASKED)
    assert has_alpha and np.all((packed.view(np.uint32) >> 24) == 200)
    r2 = ds.load_colmap(root, invert_masks=True)
    v = next(v for v in r2.train + r2.eval if v.mask_path)
    packed, _ = v.load_packed(ds.ALPHA_MASKED)
    assert v.invert_mask and np.all((packed.view(np.uint32) >> 24) == 55)


# ---------------------------------------------------------------------------------------------- nerfstudio json
def _write_nerfstudio(root, scene_extra=None, frames=None, name="transforms.json", images=("a.png", "b.png", "c.png")):
    import json
    from PIL import Image
    os.makedirs(os.path.join(root, "images"), exist_ok=True)
    for nm in images:
        Image.new("RGB", (IMG_W, IMG_H), (10, 20, 30)).save(os.path.join(root, "images", nm))
    ident = [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]]
    if frames is None:
        frames = [{"file_path": f"images/{nm}", "transform_matrix": [[1, 0, 0, float(i)], [0, 1, 0, 2.0], [0, 0, 1, 3.0], [0, 0, 0, 1]]}
                  for i, nm in enumerate(images)]
    scene = {"camera_angle_x": 1.0, "frames": frames}
    scene.update(scene_extra or {})
    json.dump(scene, open(os.path.join(root, name), "w"))
    return ident


def test_opengl_c2w_to_pose():
    """formats/mod.rs:122-131: the identity OpenGL pose looks down -Z with +Y up; in brush's convention (+Y down, +Z forward)
    that is a half turn about X.  Translation passes through; a uniform scale is divided out."""
    pos, rot = ds.opengl_c2w_to_pose(np.eye(4))
    assert pos == (0.0, 0.0, 0.0)
    np.testing.assert_allclose(np.abs(rot), [1, 0, 0, 0], atol=1e-12)
    m = np.eye(4); m[:3, :3] *= 2.5; m[:3, 3] = [1, 2, 3]
    pos, rot = ds.opengl_c2w_to_pose(m)
    assert pos == (1.0, 2.0, 3.0)
    np.testing.assert_allclose(np.abs(rot), [1, 0, 0, 0], atol=1e-12)
    # a camera yawed 90 degrees about the world up axis: its forward (-Z in OpenGL) points along -X
    c, s_ = 0.0, 1.0
    yaw = np.array([[c, 0, s_, 0], [0, 1, 0, 0], [-s_, 0, c, 0], [0, 0, 0, 1]], float)
    _, rot = ds.opengl_c2w_to_pose(yaw)
    R = cm._mat3_from_quat_xyzw(rot)                 # columns = the camera's axes in the world
    np.testing.assert_allclose(np.asarray(R)[2], [-1, 0, 0], atol=1e-6)     # brush forward (+Z local)
    np.testing.assert_allclose(np.asarray(R)[1], [0, -1, 0], atol=1e-6)     # brush +Y is down


def test_loads_nerfstudio_json(tmp_path):
    root = str(tmp_path)
    _write_nerfstudio(root)
    r = ds.load_dataset(root, eval_split_every=2)
    assert [os.path.basename(v.image_path) for v in r.eval] == ["a.png", "c.png"] and len(r.train) == 1 and not r.warnings
    v = r.train[0]
    assert v.camera.position == (1.0, 2.0, 3.0) and v.camera.camera_model == cm.PINHOLE
    assert abs(v.camera.fov_x - 1.0) < 1e-12
    focal = cm.fov_to_focal(1.0, IMG_W)                                  # the missing fov follows from the same focal length
    assert abs(v.camera.fov_y - cm.focal_to_fov(focal, IMG_H)) < 1e-12
    assert v.camera.center_uv == (0.5, 0.5) and r.init_splat is None


def test_nerfstudio_intrinsics_overrides_and_models(tmp_path):
    root = str(tmp_path)
    mat = [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]]
    frames = [{"file_path": "images/a.png", "transform_matrix": mat},
              {"file_path": "images/b", "transform_matrix": mat, "fl_x": 90.0, "fl_y": 85.0, "cx": 16.0, "cy": 12.0, "w": 64, "h": 48,
               "camera_model": "OPENCV_FISHEYE", "k1": 0.01, "k3": 0.003},
              {"file_path": "images/missing.png", "transform_matrix": mat}]
    _write_nerfstudio(root, {"camera_angle_x": None, "fl_x": 80.0, "fl_y": 70.0, "camera_model": "OPENCV", "k1": 0.1, "p2": -0.02,
                             "cx": 30.0, "cy": 20.0}, frames)
    r = ds.load_nerfstudio(root)
    assert len(r.train) == 2 and r.warnings == ["Skipped 'images/missing.png': image file not found"]
    a, b = r.train
    assert a.camera.camera_model == cm.RADIAL_TANGENTIAL_8
    np.testing.assert_allclose(a.camera.model_params, [0.1, 0, 0, 0, 0, 0, 0, -0.02], rtol=1e-7)
    assert abs(a.camera.fov_x - cm.focal_to_fov(80.0, IMG_W, a.camera.camera_model, a.camera.model_params)) < 1e-12
    assert abs(a.camera.fov_y - cm.focal_to_fov(70.0, IMG_H, a.camera.camera_model, a.camera.model_params)) < 1e-12
    np.testing.assert_allclose(a.camera.center_uv, (30.0 / IMG_W, 20.0 / IMG_H), rtol=1e-6)
    # frame b: its own model / focal / centre; k2, k4 fall back to the file level (absent -> 0); extension-less path -> png
    assert b.image_path.endswith("images/b.png") and b.camera.camera_model == cm.KANNALA_BRANDT_4
    np.testing.assert_allclose(b.camera.model_params, [0.01, 0, 0.003, 0], rtol=1e-6)
    assert abs(b.camera.fov_x - cm.focal_to_fov(90.0, 64, b.camera.camera_model, b.camera.model_params)) < 1e-12
    np.testing.assert_allclose(b.camera.center_uv, (0.25, 0.25), rtol=1e-6)


def test_nerfstudio_val_file_ply_and_errors(tmp_path):
    import json
    root = str(tmp_path)
    _write_nerfstudio(root, {"ply_file_path": "sparse.ply"}, name="transforms_train.json")
    _write_nerfstudio(root, name="transforms_val.json", images=("a.png",))
    xyz = np.array([[0, 0, 1], [1, 0, 2], [0, 1, 3], [1, 1, 4]], np.float32)
    from brush_b200 import ply
    open(os.path.join(root, "sparse.ply"), "wb").write(ply.splat_to_ply(
        np.concatenate([xyz, np.tile([1, 0, 0, 0], (4, 1)), np.full((4, 3), -3.0)], 1).astype(np.float32),
        np.zeros((4, 1, 3), np.float32), np.zeros(4, np.float32)))
    r = ds.load_nerfstudio(root, eval_split_every=2)
    assert len(r.train) == 3 and len(r.eval) == 1          # a val file exists: no training view is diverted to eval
    np.testing.assert_array_equal(r.init_splat.means, xyz)
    # load_dataset: an init.ply anywhere overrides the format's own points (formats/mod.rs:88-103)
    os.makedirs(os.path.join(root, "extra"))
    open(os.path.join(root, "extra", "init.ply"), "wb").write(ply.splat_to_ply(
        np.concatenate([xyz[:2] + 10, np.tile([1, 0, 0, 0], (2, 1)), np.full((2, 3), -3.0)], 1).astype(np.float32),
        np.zeros((2, 1, 3), np.float32), np.zeros(2, np.float32)))
    r = ds.load_dataset(root)
    np.testing.assert_array_equal(r.init_splat.means, xyz[:2] + 10)
    # unsupported model, no focal at all, a 3x4 matrix, nothing recognisable, no usable view
    bad = str(tmp_path / "bad"); os.makedirs(bad)
    _write_nerfstudio(bad, {"camera_model": "EQUIRECTANGULAR"})
    with pytest.raises(ValueError, match="Unsupported nerfstudio camera_model"):
        ds.load_dataset(bad)
    _write_nerfstudio(bad, {"camera_angle_x": None})
    with pytest.raises(ValueError, match="Must have some kind of focal length"):
        ds.load_dataset(bad)
    _write_nerfstudio(bad, frames=[{"file_path": "images/a.png", "transform_matrix": [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]]}])
    with pytest.raises(ValueError, match="12-element transform_matrix"):
        ds.load_dataset(bad)
    _write_nerfstudio(bad, frames=[{"file_path": "images/nope.png", "transform_matrix": np.eye(4).tolist()}])
    with pytest.raises(ValueError, match="no usable training views"):
        ds.load_dataset(bad)
    empty = str(tmp_path / "empty"); os.makedirs(empty)
    with pytest.raises(ValueError, match="Format not recognized"):
        ds.load_dataset(empty)


def test_find_image_by_name_never_returns_a_mask():
    """formats/mod.rs:112-120."""
    files = ["images/sub/img.png", "masks/sub/img.png", "images_2/sub/img.png", "other/IMG.PNG"]
    assert ds.find_image_by_name(files, "img.png") == "images/sub/img.png"            # min over the non-mask candidates
    assert ds.find_image_by_name(files, "sub/img.png") == "images/sub/img.png"
    assert ds.find_image_by_name(["masks/img.png"], "img.png") is None
    assert ds.find_image_by_name(files, "mg.png") is None                             # component boundary
    assert ds.find_image_by_name(["other/IMG.PNG"], "img.png") == "other/IMG.PNG"     # case-insensitive keys


# ---------------------------------------------------------------------------------------------- RealityCapture csv
RC_HEADER = "#name,x,y,alt,heading,pitch,roll,f,px,py,k1,k2,k3,k4,t1,t2"
RC_ROW = ("frame_00001.jpeg,44.5876747664166,138.823621534044,6.821916401534405,170.3483067926429,85.18637269312288,"
          "-22.74995074830745,14.2390682243052,-9.482184385774318e-004,-2.446553068050568e-004,3.114799768048152e-003,"
          "4.026391718074555e-003,-1.795976992379612e-003,0,0,0")


def test_realitycapture_header_and_rows():
    """realitycapture.rs:229-297 restated."""
    header = ds.rc_parse_header(RC_HEADER)
    assert header["name"] == 0 and header["alt"] == 3 and header["t2"] == 15
    assert ds.rc_parse_header("a,b,c") is None
    cam = ds.rc_row_to_camera(RC_ROW.split(","), header, 3840, 2880)
    assert cam.is_valid()
    np.testing.assert_allclose(cam.position, (44.587674, 138.82362, 6.821916), atol=1e-3)   # the basis swap does not move the camera
    np.testing.assert_allclose(cam.center_uv, (0.5, 0.5), atol=1e-2)
    assert cam.fov_x > cam.fov_y > 0.0 and 1.0 < cam.fov_x < 2.0                          # f = 14.24 mm (35 mm equiv.), 4:3
    assert cam.camera_model == cm.RADIAL_TANGENTIAL_8
    # a customised template without principal point and distortion columns: a centred pinhole
    h2 = ds.rc_parse_header("#name,x,y,alt,heading,pitch,roll,f")
    cam = ds.rc_row_to_camera("img.png,1,2,3,10,20,30,20.0".split(","), h2, 1920, 1080)
    assert cam.is_valid() and cam.camera_model == cm.PINHOLE and cam.center_uv == (0.5, 0.5)
    np.testing.assert_allclose(cam.position, (1.0, 2.0, 3.0), atol=1e-6)
    assert ds.rc_build_camera_model(0, 0, 0, 0, 0) == (cm.PINHOLE, ())
    assert ds.rc_build_camera_model(1.0, 2.0, 3.0, 4.0, 5.0) == (cm.RADIAL_TANGENTIAL_8, (1.0, 2.0, 3.0, 0.0, 0.0, 0.0, 4.0, 5.0))


def test_realitycapture_orientation():
    """heading = pitch = roll = 0 is a camera looking straight down (-Z of the OpenGL basis is world -Z); pitch 90 levels
    it to look along +Y; heading then turns it clockwise seen from above (yaw(-heading) about Z)."""
    h = ds.rc_parse_header("#name,x,y,alt,heading,pitch,roll,f")
    fwd = lambda head, pitch: np.asarray(cm._mat3_from_quat_xyzw(
        ds.rc_row_to_camera(f"i.png,0,0,0,{head},{pitch},0,20".split(","), h, 100, 100).rotation))[2]
    np.testing.assert_allclose(fwd(0, 0), [0, 0, -1], atol=1e-6)
    np.testing.assert_allclose(fwd(0, 90), [0, 1, 0], atol=1e-6)
    np.testing.assert_allclose(fwd(90, 90), [1, 0, 0], atol=1e-6)


def test_loads_realitycapture_csv(tmp_path):
    from PIL import Image
    root = str(tmp_path)
    os.makedirs(os.path.join(root, "imgs"))
    for nm in ("a.png", "b.png", "c.png"):
        Image.new("RGB", (80, 60), (1, 2, 3)).save(os.path.join(root, "imgs", nm))
    open(os.path.join(root, "notes.csv"), "w").write("a,b,c\n1,2,3\n")
    open(os.path.join(root, "cams.csv"), "w").write(
        RC_HEADER + "\n\n" + "a.png,1,2,3,10,80,0,20,0,0,0,0,0,0,0,0\n" + "b.png,4,5,6,20,85,1,20,0.01,0,0.01,0,0,0.5,0,0\n"
        + "missing.png,0,0,0,0,0,0,20,0,0,0,0,0,0,0,0\n" + "c.png,7,8,9,30,90,2,20,0,0,0,0,0,0,0,0\n")
    r = ds.load_dataset(root, eval_split_every=3)
    assert [os.path.basename(v.image_path) for v in r.eval] == ["a.png"] and [os.path.basename(v.image_path) for v in r.train] == ["b.png", "c.png"]
    assert r.init_splat is None
    assert r.warnings == ["RealityCapture brown4 radial term (k4) isn't supported; approximating with brown3",
                          "Skipped 'missing.png': image file not found"]
    b = r.train[0].camera
    assert b.camera_model == cm.RADIAL_TANGENTIAL_8 and b.position == (4.0, 5.0, 6.0)
    np.testing.assert_allclose(b.center_uv, ((0.01 * 80 + 40) / 80, 0.5), rtol=1e-6)
    assert abs(r.eval[0].camera.fov_x - cm.focal_to_fov(20.0 * 80 / 36.0, 80)) < 1e-12
