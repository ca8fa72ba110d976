"""tc vs fp32 nearest-code agreement over a grid of shapes (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dct_autoencoder_b200.vector_quantize import nearest_code
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(0)
for T, C, d in [(256, 256, 64), (256, 256, 128), (256, 256, 256), (256, 512, 256), (256, 2048, 256), (256, 8192, 256),
                (3000, 8192, 256), (3000, 8192, 64), (3000, 512, 256), (1000, 8192, 128), (777, 4096, 192)]:
    x = torch.randn(T, d, device=dev, generator=g) * 3
    e = torch.randn(C, d, device=dev, generator=g)
    i_tc, _ = nearest_code(x, e, impl="tc")
    i_32, _ = nearest_code(x, e, impl="fp32")
    torch.cuda.synchronize()
    bad = (i_tc != i_32)
    print(T, C, d, "mismatch", int(bad.sum()), "first bad tokens", bad.nonzero()[:8].flatten().tolist(),
          "tc", i_tc[bad][:6].tolist(), "ref", i_32[bad][:6].tolist(), flush=True)
