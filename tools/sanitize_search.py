import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
from nano_hevc_b200 import batched
import oracle as O
rng = np.random.default_rng(5)
for n in (8, 16, 32):
    H, W = 5 * n + 3, 8 * ((21 * n) // 8) + 8
    src = rng.integers(0, 256, (H, W)).astype(np.int16)
    src[H // 2:] = np.clip(60 + np.arange(W)[None] // 3 + rng.integers(-3, 4, (H - H // 2, W)), 0, 255)
    r = batched.encode_frame(torch.from_numpy(src).cuda(), n, cost="satd", qp=27)
    w = O.encode_frame(src, n, cost="satd", qp=27, recon_neighbours=False)
    torch.cuda.synchronize()
    assert np.array_equal(r.modes.cpu().numpy(), w["modes"]) and np.array_equal(r.costs.cpu().numpy(), w["costs"]), n
print("ok")
