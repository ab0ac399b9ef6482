import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fusion4landslide_b200 import ops, _lib
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for D in (32, 64):
    N, M = 3000, 5000
    A = torch.nn.functional.normalize(torch.randn(N, D, device=dev, generator=g), dim=1)
    B = torch.nn.functional.normalize(torch.randn(M, D, device=dev, generator=g), dim=1)
    A[:1000] = torch.nn.functional.normalize(B[:1000] + 0.1 * torch.randn(1000, D, device=dev, generator=g), dim=1)
    idx, d2 = ops.desc_nn(A, B, algo="tensor")
    torch.cuda.synchronize()
    ref = torch.cdist(A.double(), B.double()).argmin(1)
    print(D, "match", float((idx.long() == ref).float().mean()), idx[:8].tolist(), ref[:8].tolist(), d2[:4].tolist())
