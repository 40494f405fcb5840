import sys, importlib, numpy as np, time
sys.path.insert(0,'.')
capi = importlib.import_module('teaser-plusplus_b200.capi'); synth = importlib.import_module('teaser-plusplus_b200.synth')
rng=np.random.default_rng(1889)
src=np.transpose(synth.read_ply_vertices('tests/golden/bun_zipper_res3.ply').astype(np.float64)); N=src.shape[1]
T=np.array([[9.96926560e-01,6.68735757e-02,-4.06664421e-02,-1.15576939e-01],[-6.61289946e-02,9.97617877e-01,1.94008687e-02,-3.87705398e-02],[4.18675510e-02,-1.66517807e-02,9.98977765e-01,1.14874890e-01],[0,0,0,1]])
dst=T[:3,:3]@src+T[:3,3:4]; dst+=(rng.random((3,N))-0.5)*2*0.05
oi=rng.integers(1700,size=1700)
for i in range(oi.size):
    shift=5+rng.random((3,1))*5; dst[:,oi[i]]+=shift.squeeze()
ctx=capi.Context(0)
S=np.ascontiguousarray(src.T); D=np.ascontiguousarray(dst.T)
for lim in (0.5, 5.0):
    for mode in (0,1):
        p=capi.default_params(noise_bound=0.05,cbar2=1.0,estimate_scaling=0,rotation_cost_threshold=1e-12,max_clique_time_limit=lim,inlier_selection_mode=mode)
        t=time.time(); r=ctx.solve(S,D,p); dt=time.time()-t
        print('limit',lim,'mode',mode,'clique',len(r['clique']),'proven',r['proven'],'time',round(dt,2),'rot err',synth.angular_error(T[:3,:3],r['R']),'t err',np.linalg.norm(T[:3,3]-r['t']), 'stage', [round(x,1) for x in r['stage_ms'][:4]], flush=True)
print(ctx.debug_counters())
print("proven flag", int(r["sol"].clique_proven_optimal))
