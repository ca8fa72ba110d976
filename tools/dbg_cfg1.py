import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import dcta_oracle as O
import dct_autoencoder_b200 as D
g = np.load(os.path.join(ROOT, "tests/golden/config1.npz"))
ims = torch.from_numpy(g["images"]).float() / 255
U = D.util
print("fold_ok 256/252:", U.fold_ok(256, 256, 252, 252), "tc_forward_ok", U.tc_forward_ok(256, 256))
for i in (0, 3, 11):
    x = ims[i:i+1].cuda()
    ref = O.transform_image_in(ims[i].numpy())[:, :252, :252]
    ref64 = __import__("scipy.fft").fft.dctn(O.rgb_to_ipt(ims[i].numpy()).astype(np.float64), type=2, norm="ortho", axes=(-2,-1))[:, :252, :252]
    for impl in ("tc", "tc_plain", "fp32"):
        fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072, dct_impl=impl)
        tiles = fe._token_grid(x)   # (1, th, tw, c, 196)
        th, tw = tiles.shape[1:3]
        plane = tiles[0].reshape(th, tw, 3, 14, 14).permute(2, 0, 3, 1, 4).reshape(3, th*14, tw*14).cpu().numpy()
        err = np.abs(plane - ref64)
        j = np.unravel_index(err.argmax(), err.shape)
        print(i, impl, "max err %.3e (%.2e of max|Y| %.1f) at" % (err.max(), err.max()/np.abs(ref64).max(), np.abs(ref64).max()), j, "value", ref64[j],
              "| ipt err", np.abs(U.rgb_to_ipt(x)[0].cpu().numpy() - O.rgb_to_ipt(ims[i].numpy())).max())
