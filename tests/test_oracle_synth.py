"""CPU: the oracle's on-the-fly synthetic-map loop (oracle_render_synth, used for parity bands at the BASELINE
map sizes) is the same function as oracle_render over the materialised maps."""
import numpy as np
import pytest

import bench


@pytest.mark.parametrize("workload", ["spherical1080", "ortho4k", "flythrough4k"])
def test_render_synth_equals_render_over_arrays(oracle, workload):
    wl = dict(bench.WORKLOADS[workload])
    wl["log2n"] = 9                       # same cameras, scaled to a 512^2 map (the camera scales with the extent)
    W, H = 160, 90
    wl["W"], wl["H"] = W, H
    c = bench.camera(wl, 7)
    if wl["projection"] == 3:             # the orthographic bench camera has a fixed height meant for an 8192^2 map
        c["pos"] = (c["pos"][0], c["pos"][1], 30.0 * 512 / 8192)
    fr = oracle.make_frame(projection=wl["projection"], width=W, height=H, grid_width=bench.GRID_WIDTH,
                           step_dist=wl["step_dist"], min_height=bench.MIN_HEIGHT, max_height=bench.MAX_HEIGHT, **c)
    hm, cm = oracle.synth_maps(9, bench.SEED)
    heights = oracle.update_heightmap(hm, (0.299, 0.587, 0.114), bench.MIN_HEIGHT, bench.MAX_HEIGHT)
    want, wsteps, wst = oracle.render(fr, heights, cm)
    got, gsteps, gst = oracle.render_synth(fr, 9, bench.SEED)
    assert np.array_equal(got, want) and np.array_equal(gsteps, wsteps)
    assert (gst.steps, gst.surf_hits, gst.box_hits) == (wst.steps, wst.surf_hits, wst.box_hits)
    assert wst.surf_hits > 0
    # a row band writes only its rows
    band, bsteps, _ = oracle.render_synth(fr, 9, bench.SEED, rows=(40, 44))
    assert np.array_equal(band[40:44], want[40:44]) and not band[:40].any() and not band[44:].any()
    assert np.array_equal(bsteps[40:44], wsteps[40:44])
