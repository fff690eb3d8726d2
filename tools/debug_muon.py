import sys, numpy as np
sys.path.insert(0,'.')
from tests.gpu_util import shower, primaries
from tests.parity import oracle_showers
sh = shower('graphite', 0.030, seed=11)
prims = primaries(13, 20.0, 1)
b = sh.generate_showers(prims, first_shower_id=1000)
ref = oracle_showers(prims, 'graphite', 0.030, 11, first_shower_id=1000)[0]
h = b.to_host(); order, offs = b.reference_order()
bad = 0
for k,(s,q) in enumerate(zip(order, ref)):
    d = np.max(np.abs(np.asarray(q.rf)-h['rf'][s]))
    if d > 1e-6:
        bad += 1
        if bad < 6: print(k, q.PID, q.process, 'p0', q.p0, 'pf', q.pf, h['pf'][s], 'r0', q.r0, h['r0'][s], 'rf', q.rf, h['rf'][s], 'nsub', q.nsub, h['nsub'][s])
print('bad', bad, 'of', len(ref))
