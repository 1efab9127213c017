import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import geodesic_raytracing_b200 as grt
content = grt.Content(grt.CONTENT_DIR)
metric = content.load_by_name("kerr_schild")
ctx = grt.Context(0)
ctx.set_metric(metric)
ctx.set_background(grt.synthetic_background(512, 256))
ctx.set_camera()
ctx.set_features(adaptive_sampling=0, use_prepass=0)
w, h = 256, 144
res = {}
for passes, thr in [(3, 0), (3, 16), (3, 31), (0, 31)]:
    ctx.set_tuning(passes, thr)
    img, st = ctx.render(w, h)
    rays = ctx.read_rays().copy()
    res[(passes, thr)] = (rays, st.trial_steps, st.accepted_steps, st.rays_integrated, st.rays_finished)
    print(passes, thr, 'steps', st.trial_steps, st.accepted_steps, 'rays', st.rays_integrated, st.rays_finished, 'term hist', np.bincount(rays['terminated'], minlength=5))
a = res[(3, 0)][0]
for k, v in res.items():
    b = v[0]
    diff = (a['terminated'] != b['terminated'])
    fin = (a['terminated'] == 1) & (b['terminated'] == 1)
    pd = np.abs(a['position'][fin] - b['position'][fin]).max() if fin.any() else 0
    print(k, 'terminated mismatches', diff.sum(), 'max pos diff among finished', pd, 'dlambda diff', np.abs(a['running_dlambda_dnew'] - b['running_dlambda_dnew']).max())
    idx = np.where(diff)[0][:5]
    for i in idx:
        print('   slot', i, 'a', a[i]['terminated'], a[i]['position'], 'b', b[i]['terminated'], b[i]['position'])
    pdiff = np.abs(a['position'] - b['position']).max(axis=1)
    j = np.where((pdiff > 0) & fin)[0][:5]
    for i in j:
        print('   pos diff slot', i, a[i]['position'], b[i]['position'], a[i]['velocity'], b[i]['velocity'])
