import os, sys, subprocess, json
# each cell size in its own process (the knob is read once)
if len(sys.argv) > 1:
    import numpy as np, torch
    sys.path.insert(0, '.')
    from avatarcap_b200 import synth, pipeline
    from avatarcap_b200.engine import Engine
    eng = Engine(); dev = eng.device
    body = synth.SynthBody(); fr = synth.make_frame(body)
    res = (256, 256, 256)
    grid = eng.make_grid(fr['cano_bounds'], res)
    cv = torch.from_numpy(fr['cano_smpl_v']).to(dev); sw = torch.from_numpy(fr['smpl_skinning_weights']).to(dev); jm = torch.from_numpy(fr['cano2live_jnt_mats']).to(dev)
    vol = torch.from_numpy(synth.body_sdf(grid.cpu().numpy(), synth.cano_pose()).reshape(res)).to(dev)
    out = {}
    for name, off in (('surface', 0.0), ('off2cm', 0.02), ('off5cm', 0.05)):
        v, f, n = eng.extract_mesh(vol + off, fr['cano_bounds'], 0.0)
        def t(fn, reps=5):
            fn(); torch.cuda.synchronize(); ts = []
            for _ in range(reps):
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
            return float(np.median(ts))
        out[name] = (int(v.shape[0]), round(t(lambda: eng.skin_mesh(v, n, cv, sw, jm)), 3))
    out['near_flag_ms'] = round(t(lambda: eng.near_flag(grid, cv, 0.1), 3), 3)
    print(json.dumps({os.environ.get('AVC_KNN_CELL', 'default'): out}))
else:
    for h in ('0.03', '0.04', '0.05', '0.06', '0.08', '0.10'):
        subprocess.run([sys.executable, __file__, 'x'], env=dict(os.environ, AVC_KNN_CELL=h))
