import sys, time, numpy as np
sys.path.insert(0, '.')
from petite_b200.train import Trainer
from petite_b200 import tables as tb
from petite_b200.shower import Shower, process_code
DATA='data/'
P = sys.argv[1] if len(sys.argv) > 1 else 'Brem'
xs = np.load(DATA + "sm_xsec.npz")[f"{P}/graphite"]
rows=[30,60,80,99]; E=xs[rows,0]
sh = Shower(DATA, "graphite", 0.010, seed=3)
shipped = sh._maps[P]
mf_old, sg_old = sh.find_max(P, n_trials=100, seed=9)
print('shipped eff', (sg_old/(300*mf_old))[rows])
for nitn, npts, alpha in [(12,400_000,0.5),(30,1_000_000,0.5),(30,1_000_000,1.0),(40,2_000_000,0.75)]:
    tr = Trainer(); t=time.time()
    grids, ninc, I = tr.train(P, E, nitn=nitn, n_points=npts, alpha=alpha)
    dt=time.time()-t
    ms = tb.MapSet(P, E, ninc, grids, np.ones(len(E)), shipped.neval, shipped.Eg_min, shipped.Ee_min)
    sh._upload_maps(process_code[P], ms); sh._maps[P] = ms
    mf, sg = sh.find_max(P, n_trials=100, seed=9)
    print(nitn, npts, alpha, 'time %.1f'%dt, 'sigma ratio', sg/xs[rows,1], 'eff', sg/(300*mf))
