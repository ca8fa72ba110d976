"""A small end-to-end workload for compute-sanitizer: every kernel family of the fused and staged paths at tiny sizes.

    compute-sanitizer --tool memcheck python tools/sanitize_small.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import dct_autoencoder_b200 as D  # noqa: E402

torch.manual_seed(0)
dev = torch.device("cuda", 0)
for (h, w, b, msl) in [(128, 112, 3, 3072), (90, 101, 2, 200)]:
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, msl)
    pn = D.PatchNorm(32, 32, 14, 3).to(dev)
    lfq = D.LFQ(codebook_size=2 ** 14, num_codebooks=14).to(dev).eval()
    pipe = D.TransformPipeline(fe, pn, lfq)
    pipe.fit_norm(torch.rand(2, 3, h, w, device=dev))
    x = torch.rand(b, 3, h, w, device=dev)
    rec, codes = pipe.roundtrip(x)                       # fused: fold_codes, bit planes, generated-operand GEMM
    rec_s, codes_s = pipe.roundtrip(x, fused=False)      # staged modules
    assert torch.equal(codes, codes_s) and torch.equal(rec, rec_s)
    batch, codes2 = pipe.encode_codes(x)
    rec2 = pipe.decode_codes(batch, codes2, out_dtype=torch.uint8)
lfq16 = D.LFQ(dim=196, codebook_size=8192, num_codebooks=16).to(dev).eval()
with torch.no_grad():
    lfq16(torch.randn(2, 40, 196, device=dev), torch.ones(2, 40, dtype=torch.bool, device=dev))
vq = D.VectorQuantize(dim=64, codebook_size=128).to(dev)
vq.eval()
vq(torch.randn(2, 70, 64, device=dev))
vq.train()
vq(torch.randn(2, 70, 64, device=dev), mask=torch.rand(2, 70, device=dev) > 0.2)
from dct_autoencoder_b200.dataset import resize_antialias  # noqa: E402
resize_antialias(torch.rand(3, 97, 131, device=dev), (31, 40))
torch.cuda.synchronize()
print("sanitize_small: ok")
