import sys, time, numpy as np
sys.path.insert(0,'.')
import __graft_entry__ as g
t=time.time(); g.smoke(); print('smoke s', time.time()-t)
from petite_b200 import Particle
from petite_b200.shower import Shower
from petite_b200.constants import m_electron
s = Shower('data/', 'graphite', 0.010, seed=1)
E=10.0
for n in (1000, 10000):
    prims=[Particle([E,0,0,np.sqrt(E**2-m_electron**2)],[0,0,0],{'PID':11,'ID':1,'mass':m_electron}) for _ in range(n)]
    import torch
    for rep in range(2):
        torch.cuda.synchronize(); t=time.time()
        b=s.generate_showers(prims, capacity=n*1200)
        torch.cuda.synchronize(); dt=time.time()-t
        print(n, 'showers', dt, 's', b.counters, 'showers/s', n/dt, 'steps/s', b.counters['n_steps']/dt)
