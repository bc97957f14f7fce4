"""Static check of the shipped library (no GPU): the sm_100a instruction selection DESIGN.md claims is in the SASS of
brush_b200/libbrush_b200.so -- TMA bulk copies (UBLKCP) in the cull kernel, TMA tile::gather4 (UTMALDG) and packed FP32
(FFMA2 / FMUL2 / FADD2) in the production blend kernels, no per-lane LDGSTS staging left in them, no tensor-core
instruction anywhere (the path has no dense contraction), and no register spills in the hot kernels."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def sass():
    if not os.path.exists(CUOBJDUMP):
        pytest.skip("cuobjdump not available")
    from brush_b200 import build
    lib = build.build()
    out = subprocess.run([CUOBJDUMP, "-sass", lib], capture_output=True, text=True, timeout=600).stdout
    hist, cur, arch = collections.OrderedDict(), None, set()
    for line in out.splitlines():
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            arch.add(m.group(1))
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            hist[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            hist[cur][m.group(1).split(".")[0]] += 1
    assert hist, "no SASS found in the library"
    return hist, arch


def _kernels(hist, needle):
    ks = {k: v for k, v in hist.items() if needle in k}
    assert ks, f"no kernel matching {needle}"
    return ks


def test_built_for_sm_100a_only(sass):
    _, arch = sass
    assert arch == {"sm_100a"}, arch


def test_cull_kernel_stages_rows_with_tma_bulk_copies(sass):
    hist, _ = sass
    for name, h in _kernels(hist, "project_cull_kernel").items():
        assert h["UBLKCP"] >= 1, (name, dict(h))


def test_blend_kernels_use_tma_gather_and_packed_fp32(sass):
    hist, _ = sass
    for needle in ("blend_fwd_kernel", "blend_bwd_kernel"):
        for name, h in _kernels(hist, needle).items():
            assert h["UTMALDG"] >= 1, (name, "no TMA tile::gather4")
            assert h["FFMA2"] >= 1 and h["FMUL2"] >= 1 and h["FADD2"] >= 1, (name, "no packed FP32")
            assert h["LDGSTS"] == 0, (name, "per-lane cp.async staging is back")
            assert h["MUFU"] >= 2, name


def test_loss_kernel_runs_its_windows_on_packed_fp32(sass):
    hist, _ = sass
    for name, h in _kernels(hist, "image_loss_fused_kernel").items():
        assert h["FFMA2"] >= 500, (name, h["FFMA2"])


def test_no_tensor_core_instructions(sass):
    hist, _ = sass
    for name, h in hist.items():
        bad = [op for op in h if op.startswith(("HMMA", "IMMA", "DMMA", "QMMA", "UTCMMA", "UTCHMMA", "HGMMA"))]
        assert not bad, (name, bad)


def test_hot_kernels_do_not_spill():
    obj = os.path.join(ROOT, "brush_b200", "csrc", "_obj")
    checked = 0
    for unit in ("blend_fwd", "blend_bwd", "sort", "update"):
        path = os.path.join(obj, unit + ".o.ptxas.txt")
        if not os.path.exists(path):
            pytest.skip("ptxas logs not present (library not built in this tree)")
        txt = open(path).read()
        for m in re.finditer(r"Function properties for (\S+)\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", txt):
            if "kernel" in m.group(1):
                assert (m.group(2), m.group(3), m.group(4)) == ("0", "0", "0"), (unit, m.group(0))
                checked += 1
    assert checked >= 6
