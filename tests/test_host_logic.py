"""CPU: host-side mirror of the reference's type-level API (hmrm_deg2rad, hmrm_camera_basis, hmrm_get_ray —
no device work) against known-answer vectors produced by the unmodified reference (tests/golden/kat.json)."""
import pytest

import helpers as H


def test_deg2rad_kat(hmrm):
    for rec in H.golden_kat()["deg2rad"]:
        assert H.same(hmrm.deg2rad(H.fx(rec["deg"])), rec["rad"])


def test_get_ray_kat(hmrm):
    n = 0
    for rec in H.golden_kat()["rays"]:
        f = hmrm.Frame()
        f.projection, f.screen_width, f.screen_height = rec["projection"], rec["W"], rec["H"]
        f.cam_pos[:] = [H.fx(v) for v in rec["pos"]]
        f.hang, f.vang, f.hfov = H.fx(rec["hang"]), H.fx(rec["vang"]), H.fx(rec["hfov"])
        f.ortho_width = H.fx(rec["ortho_width"])
        pos, d = hmrm.get_ray(f, H.fx(rec["w"]), H.fx(rec["h"]))
        assert all(H.same(v, g) for v, g in zip(pos, rec["ray_pos"])), rec
        assert all(H.same(v, g) for v, g in zip(d, rec["ray_dir"])), rec
        n += 1
    assert n >= 300


def test_camera_basis_matches_oracle(hmrm, oracle):
    import ctypes as C

    for hang, vang in [(-0.7853981633974483, 1.5707963267948966), (1.234567, 2.5), (-3.0, 0.0), (0.5, 3.141592653589793)]:
        look, up = hmrm.camera_basis(hang, vang)
        ol, ou = (C.c_double * 3)(), (C.c_double * 3)()
        oracle.lib().oracle_camera_basis(hang, vang, ol, ou)
        assert [H.bits(v) for v in look] == [H.bits(v) for v in ol]
        assert [H.bits(v) for v in up] == [H.bits(v) for v in ou]


def test_integration_patch_applies_to_the_reference_and_compiles(tmp_path):
    """oracle/make_patched_reference.py (the INTEGRATION.md edits, addressed by line number and guarded by anchors)
    still applies to the mounted reference, and the result compiles with the reference's own flags.  Running it needs
    a GPU: tests/test_gpu_patched_reference.py."""
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    ref = Path("/root/reference/main/hmap.cpp")
    if not ref.exists():
        pytest.skip("/root/reference is not mounted")
    out = tmp_path / "hmap_patched.cpp"
    subprocess.run([sys.executable, str(root / "oracle" / "make_patched_reference.py"), str(ref), str(out)], check=True)
    text = out.read_text()
    assert text.count("hmrm_render(") == 1 and "#pragma omp parallel for" not in text and "delete ip;" not in text
    res = subprocess.run(["g++", "-std=c++98", "-Wall", "-Wextra", "-Wconversion", "-fopenmp", "-fsyntax-only",
                          "-I", str(root / "oracle" / "shim"), "-I", "/root/reference/src", "-I", "/root/reference/vendor",
                          "-I", "/root/reference", "-I", str(root / "include"), str(out)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
