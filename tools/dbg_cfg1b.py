import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch, scipy.fft
import dcta_oracle as O
import dct_autoencoder_b200 as D
g = np.load(os.path.join(ROOT, "tests/golden/config1.npz"))
ims = torch.from_numpy(g["images"]).float() / 255
U = D.util
dctn = lambda a: scipy.fft.dctn(a.astype(np.float64), type=2, norm="ortho", axes=(-2, -1))
for i in (0, 11):
    x = ims[i:i+1].cuda()
    oipt = O.rgb_to_ipt(ims[i].numpy())
    ipt = U.rgb_to_ipt(x)[0].cpu().numpy()
    d = ipt.astype(np.float64) - oipt
    e = np.abs(dctn(d))
    j = np.unravel_index(e.argmax(), e.shape)
    print(i, "colour: ipt diff max %.2e mean/ch %s -> coef err %.3e at %s" % (np.abs(d).max(), d.mean((1, 2)), e.max(), j))
    # our DCT on the ORACLE's IPT
    t = torch.from_numpy(oipt)[None].cuda()
    ref = dctn(oipt)
    for name, fn in (("fp32 ffma", lambda: U.dct2_truncated(t, 252, 252)),
                     ("fold tc", lambda: None)):
        y = fn()
        if y is None:
            hi, lo, dc = U.fold_planes(t) if hasattr(U, "fold_planes") else (None, None, None)
            if hi is None:
                continue
            y = U.dct2_fwd_fold(hi, lo, dc, 252, 252)
        y = y[0].cpu().numpy()
        e = np.abs(y - ref[:, :252, :252]); j = np.unravel_index(e.argmax(), e.shape)
        print("   ", name, "on oracle ipt: max err %.3e at %s (%.2e of max)" % (e.max(), j, e.max() / np.abs(ref).max()))
print([n for n in dir(U) if "fold" in n or "dct" in n])
