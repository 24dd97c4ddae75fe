import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, hmrm_pkg, oracle_lib as O, scenes as S, helpers as H
hmrm = hmrm_pkg.load()
r = hmrm.Renderer(0)
name = sys.argv[1] if len(sys.argv) > 1 else 'spher_wide'
sc = S.SCENE_BY_NAME[name]
maps = H.load_scene_maps(sc, O)
H.configure(r, sc, maps)
res = {}
for t in (1,2):
    f = H.product_frame(hmrm, r, sc, traversal=t, flags=3)
    fb = r.render(f).copy(); si = r.step_index(f).copy(); st = r.stats()
    res[t] = (fb, si, st.steps, st.surf_hits)
    print(t, 'steps', st.steps, 'surf', st.surf_hits, 'dbg', r.debug_counters())
a, b = res[1][1], res[2][1]
d = b.astype(np.int64) - a
print('ndiff', (d!=0).sum(), 'pixdiff', (res[1][0]!=res[2][0]).any(axis=2).sum())
ys, xs = np.nonzero(d)
for y, x in list(zip(ys, xs))[:16]:
    print((y, x), 'brute', a[y, x], 'lin', b[y, x], 'col brute', res[1][0][y, x, :3], 'lin', res[2][0][y, x, :3])
print('--- rays brute hits but lin does not ---')
sel = (a >= 0) & (b < 0)
ys, xs = np.nonzero(sel)
for y, x in list(zip(ys, xs))[:20]:
    print((y, x), 'brute', a[y, x], 'lin', b[y, x])
