"""TEST INFRASTRUCTURE ONLY -- stand-in for the third-party package ``torch_dct==0.1.6``.

The reference pins ``torch_dct==0.1.6`` (/root/reference/requirements.txt:11) and
calls it from ``dct_autoencoder/util.py:9`` (import) and ``util.py:333-338``
(``dct2`` -> ``torch_dct.dct_2d``, ``idct2`` -> ``torch_dct.idct_2d``).  The package
is not vendored under /root/reference and is not installable here (no network), so
its *published* algorithm is restated below so that the unmodified reference can be
imported and run in this container to produce golden vectors:

  * DCT-II via Makhoul's N-point FFT: reorder ``v = [x[::2], reversed(x[1::2])]``,
    ``V = fft(v)``, multiply by the twiddle ``exp(-i*pi*k/(2N))``, take the real part,
    apply the orthonormal scaling (``k=0`` by ``1/(2*sqrt(N))``, ``k>0`` by
    ``1/(2*sqrt(N/2))``) and a final factor 2.
  * DCT-III (inverse) as the exact algebraic inverse of the above through ``irfft``.
  * 2-D transforms apply the 1-D transform to the last axis, then to the
    second-to-last axis.

PARITY STATUS: "parity unpinned" at the bit level -- the reference's tests hold no
golden vector for this boundary (testpatching.py:42-43 replaces the transform with
the identity) and the real 0.1.6 wheel cannot be obtained here.  The stand-in is
pinned instead to the mathematical definition (orthonormal DCT-II/III,
``scipy.fft.dctn/idctn(type=2, norm="ortho")`` in float64) by
``tests/test_oracle.py::test_standin_matches_float64_definition``.

Nothing in the product path imports this file.
"""
import math

import torch


def dct(x, norm=None):
    shape = x.shape
    n = shape[-1]
    x = x.contiguous().view(-1, n)
    v = torch.cat([x[:, ::2], x[:, 1::2].flip([1])], dim=1)
    vc = torch.view_as_real(torch.fft.fft(v, dim=1))
    k = -torch.arange(n, dtype=x.dtype, device=x.device)[None, :] * math.pi / (2 * n)
    w_r, w_i = torch.cos(k), torch.sin(k)
    out = vc[:, :, 0] * w_r - vc[:, :, 1] * w_i
    if norm == "ortho":
        out[:, 0] /= math.sqrt(n) * 2
        out[:, 1:] /= math.sqrt(n / 2) * 2
    return 2 * out.view(*shape)


def idct(X, norm=None):
    shape = X.shape
    n = shape[-1]
    xv = X.contiguous().view(-1, n) / 2
    if norm == "ortho":
        xv[:, 0] *= math.sqrt(n) * 2
        xv[:, 1:] *= math.sqrt(n / 2) * 2
    k = torch.arange(n, dtype=X.dtype, device=X.device)[None, :] * math.pi / (2 * n)
    w_r, w_i = torch.cos(k), torch.sin(k)
    vt_r = xv
    vt_i = torch.cat([xv[:, :1] * 0, -xv.flip([1])[:, :-1]], dim=1)
    v_r = vt_r * w_r - vt_i * w_i
    v_i = vt_r * w_i + vt_i * w_r
    v = torch.fft.irfft(torch.complex(v_r, v_i), n=n, dim=1)
    x = v.new_zeros(v.shape)
    x[:, ::2] += v[:, : n - (n // 2)]
    x[:, 1::2] += v.flip([1])[:, : n // 2]
    return x.view(*shape)


def dct_2d(x, norm=None):
    a = dct(x, norm=norm)
    b = dct(a.transpose(-1, -2), norm=norm)
    return b.transpose(-1, -2)


def idct_2d(X, norm=None):
    a = idct(X, norm=norm)
    b = idct(a.transpose(-1, -2), norm=norm)
    return b.transpose(-1, -2)
