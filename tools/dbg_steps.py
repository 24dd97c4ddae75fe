import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, hmrm_pkg, oracle_lib as O, scenes as S, helpers as H
hmrm = hmrm_pkg.load()
r = hmrm.Renderer(0)
sc = S.SCENE_BY_NAME['persp_basic']
maps = H.load_scene_maps(sc, O)
H.configure(r, sc, maps)
res = {}
for t in (1,2):
    f = H.product_frame(hmrm, r, sc, traversal=t, flags=3)
    fb = r.render(f).copy(); si = r.step_index(f).copy(); st = r.stats()
    res[t] = (fb, si, st.steps)
a, b = res[1][1], res[2][1]
d = b - a
print('steps', res[1][2], res[2][2], 'ndiff', (d!=0).sum())
vals, cnt = np.unique(d[d!=0], return_counts=True)
print(list(zip(vals[:20], cnt[:20])))
ys, xs = np.nonzero(d)
print(list(zip(ys[:10], xs[:10], a[ys[:10], xs[:10]], b[ys[:10], xs[:10]])))
