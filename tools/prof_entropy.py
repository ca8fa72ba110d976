import os, sys, torch
sys.path.insert(0, '/root/repo')
import dct_autoencoder_b200 as D
from dct_autoencoder_b200.util import FactorizedDistance, compute_entropy_loss
torch.cuda.set_device(0)
b, n, c, d = 64, 3072, 14, 14
x = (torch.randn(b, n, c, d, device='cuda') * 0.002).requires_grad_(True)
mask = torch.rand(b, n, device='cuda') > 0.1
for _ in range(2):
    x.grad = None
    compute_entropy_loss(FactorizedDistance(x, 1.0), mask).backward()
torch.cuda.synchronize()
