import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import pyloworder_b200 as pl, synth
A = synth.snapshots(200000, 512, 2022)
R = torch.from_numpy(np.linalg.qr(A, mode="r")).cuda().contiguous()
U, S, V = pl.math.svd(R); torch.cuda.synchronize()
