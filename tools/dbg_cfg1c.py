import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch, scipy.fft
import dcta_oracle as O
import dct_autoencoder_b200 as D
g = np.load(os.path.join(ROOT, "tests/golden/config1.npz"))
ims = torch.from_numpy(g["images"]).float() / 255
dctn = lambda a: scipy.fft.dctn(a.astype(np.float64), type=2, norm="ortho", axes=(-2, -1))
np.set_printoptions(linewidth=200, precision=2)
for impl in ("tc", "fp32"):
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072, dct_impl=impl)
    rels = []
    for i in range(13):
        x = ims[i:i+1].cuda()
        ref = dctn(D.util.rgb_to_ipt(x)[0].cpu().numpy())[:, :252, :252]     # exact DCT of OUR ipt: isolates the DCT arithmetic
        tiles = fe._token_grid(x)
        th, tw = tiles.shape[1:3]
        y = tiles[0].reshape(th, tw, 3, 14, 14).permute(2, 0, 3, 1, 4).reshape(3, th*14, tw*14).cpu().numpy().astype(np.float64)
        big = np.argsort(-np.abs(ref).reshape(-1))[1:13]      # skip the DC
        rel = ((y - ref) / ref).reshape(-1)[big]
        rels.append(rel)
        if i < 4:
            print(impl, i, "rel err of the 12 largest AC coefficients (x1e-7):", rel * 1e7, "| max abs err/max|Y| %.2e" % (np.abs(y-ref).max()/np.abs(ref).max()))
    r = np.concatenate(rels)
    print(impl, "ALL: mean rel %.3e, std %.3e, frac negative %.2f" % (r.mean(), r.std(), (r < 0).mean()))
