"""Experiment: two half-batches on two CUDA streams (two engines, two host threads) vs one full batch."""
import sys, time, threading
import numpy as np, torch
sys.path.insert(0, '.')
from petite_b200.shower import Shower
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device('cuda', 0)
def prim(k):
    p = np.tile(np.array([10.0, 0, 0, 10.0]), (k, 1)); r = np.zeros((k, 3)); w = np.ones(k); m = np.zeros(k)
    return p, r, w, m, np.full(k, 22, np.int32), np.zeros(k, np.int32)
eng = [Shower('data/', 'lead', 0.010, seed=1) for _ in range(S)]
streams = [torch.cuda.Stream() for _ in range(S)]
cap = int(N / S * 1700)
def run(i, first):
    with torch.cuda.stream(streams[i]):
        eng[i].run_arrays(*prim(N // S), capacity=cap, first_shower_id=first)
for rep in range(3):
    torch.cuda.synchronize(); t = time.time()
    th = [threading.Thread(target=run, args=(i, i * (N // S))) for i in range(S)]
    [x.start() for x in th]; [x.join() for x in th]
    torch.cuda.synchronize(); dt = time.time() - t
    print(f'{S} streams x {N//S} showers: {dt*1e3:.1f} ms  -> {N/dt:.0f} showers/s')
