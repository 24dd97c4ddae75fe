"""CPU: the C-ABI library loads and exports every symbol include/hmrm.h declares (no compute calls)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "hmrm.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hmrm_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for must in ("hmrm_create", "hmrm_set_maps", "hmrm_update_heightmap", "hmrm_render", "hmrm_render_device",
                 "hmrm_render_async", "hmrm_wait", "hmrm_get_stats", "hmrm_last_error", "hmrm_destroy"):
        assert must in syms


def test_library_exports_every_declared_symbol(hmrm):
    lib = hmrm.load_library()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/hmrm.h but not exported by libhmrm.so"


def test_binding_prototypes_cover_header(hmrm):
    from heightmap_ray_marcher_b200 import binding

    assert sorted(binding.PROTOTYPES) == declared_symbols()


def test_abi_version_and_struct_sizes(hmrm):
    lib = hmrm.load_library()
    assert lib.hmrm_abi_version() == 2
    # struct layouts mirrored in ctypes must match the C compiler's (checked through defaults)
    f = hmrm.Frame()
    lib.hmrm_frame_defaults(C.byref(f))
    assert (f.projection, f.screen_width, f.screen_height) == (1, 800, 600)      # main/hmap.cpp:107,31,32
    assert tuple(f.cam_pos) == (-5.0, 5.0, 0.0)                                   # :75
    assert (f.grid_width, f.step_dist, f.ortho_width) == (0.05, 0.25, 0.1)        # :65,68,98
    assert (f.cycle, f.cycle_period, f.traversal, f.flags) == (0, 1, 0, 0)
    assert C.sizeof(hmrm.Frame) == 128 and C.sizeof(hmrm.Stats) == 64


def test_no_cpu_fallback_without_device(hmrm):
    """On a box without a GPU the product must fail loudly instead of rendering on the CPU."""
    lib = hmrm.load_library()
    if lib.hmrm_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(hmrm.HmrmError) as e:
        hmrm.Renderer(0)
    assert "no CPU path" in str(e.value)


def test_product_does_not_reference_oracle():
    """Nothing under the product package or include/ may import, include or link the oracle."""
    pkg = ROOT / "heightmap-ray-marcher_b200"
    for path in list(pkg.rglob("*")) + list((ROOT / "include").rglob("*")):
        if path.is_file() and path.suffix in (".py", ".h", ".cuh", ".cu", ".cpp", ".hpp"):
            text = path.read_text()
            for token in ("oracle/", "oracle_", "liboracle", "hmap_oracle", "import oracle"):
                assert token not in text, f"{path} refers to the oracle ({token})"


def test_null_context_is_an_error_not_a_crash(hmrm):
    """Entry points that take a context refuse NULL with a status (no GPU needed): the C ABI never dereferences it."""
    lib = hmrm.load_library()
    assert lib.hmrm_peer_sync_mode(None) < 0
    assert lib.hmrm_wait(None) != 0
    assert lib.hmrm_peer_release(None, None, 1, None) != 0
    lib.hmrm_get_layout(None)          # a plain getter: any value, no crash
