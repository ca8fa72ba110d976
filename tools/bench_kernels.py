"""CUDA-event timings of individual entry points at the config-2 shape (for kernel work).

    python tools/bench_kernels.py [--batch 256]
"""
import argparse
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    a = ap.parse_args()
    import torch
    import dct_autoencoder_b200 as D
    U = D.util
    dev = torch.device("cuda", 0)
    B = a.batch
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    x = torch.rand(B, 3, 512, 512, device=dev, generator=g)

    def t(name, fn, reps=5):
        fn()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"{name:48s} {statistics.median(ts) * 1e3:9.1f} us")
        return r

    hi, lo, dc = t("rgb_to_ipt_fold", lambda: U.rgb_to_ipt_fold(x))
    t("fwd fold -> token grid + maxabs", lambda: U.dct2_fwd_fold(hi, lo, dc, 448, 448, tile_p=14, channels=3, with_maxabs=True))
    t("fwd fold -> token grid", lambda: U.dct2_fwd_fold(hi, lo, dc, 448, 448, tile_p=14, channels=3))
    y = t("fwd fold -> planes", lambda: U.dct2_fwd_fold(hi, lo, dc, 448, 448, out_shape=(B, 3)))
    pn = D.PatchNorm(32, 32, 14, 3).to(dev)
    t("fwd fold -> code grid + maxabs", lambda: U.dct2_fwd_fold_codes(hi, lo, dc, 448, 448, 14, 3, pn))
    del hi, lo
    t("idct fold (fold_coef + 2 passes + unfold planes)", lambda: U.idct2_truncated_fold(y, 512, 512))


if __name__ == "__main__":
    main()
