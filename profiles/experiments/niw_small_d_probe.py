import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from distributions_b200 import capi, synth
ctx = capi.Context(0)
for (d, G, N) in ((3, 64, 1_000_000), (8, 256, 500_000)):
    w = synth.niw(134, G, N, d=d)
    f = ctx.feature(capi.NIW).update_all(w)
    col = torch.from_numpy(np.ascontiguousarray(w["values"], dtype=np.float32)).cuda()
    u = torch.from_numpy(w["u"]).cuda()
    prior = torch.empty(G, device="cuda"); ctx.prior_pitman_yor(synth.PY_ALPHA, synth.PY_D, w["sizes"], prior)
    assign = torch.empty(N, device="cuda", dtype=torch.int32)
    for _ in range(3):
        ctx.score_sample_batch([f], [col], N, prior, u, assign)
    torch.cuda.synchronize()
