"""CPU tests of the host side: the C-ABI library loads and exports what include/dcta.h declares,
host-side packing logic matches the oracle, the product path refuses to run without CUDA, and the
multi-rank statistic protocol (gloo, world_size 2) reproduces the sequential reference rule."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import dcta_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def D():
    import dct_autoencoder_b200 as d
    return d


def header_symbols():
    src = open(os.path.join(ROOT, "include", "dcta.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(dcta_[a-z0-9_]+)\s*\(", src))


def test_library_exports_every_declared_symbol(D):
    lib = D._lib.load()      # raises if libdcta.so is missing or a symbol cannot be bound
    syms = header_symbols()
    assert len(syms) >= 28
    for s in syms:
        assert hasattr(lib, s), s
    assert set(D._lib.SIGNATURES) == syms
    assert lib.dcta_abi_version() == D._lib.ABI_VERSION
    assert lib.dcta_compiled_arch() == 100
    out = subprocess.run(["nm", "-D", D._lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (dcta_[a-z0-9_]+)", out))
    assert syms <= exported


def test_invalid_arguments_return_error_codes_without_a_gpu(D):
    lib = D._lib.load()
    rc = lib.dcta_dct2_fwd(None, None, None, None, None, 1, 8, 8, 8, 8, 0, 1, None)
    assert rc == -1 and "null pointer" in D._lib.last_error()
    rc = lib.dcta_sort_tokens(1, 1, 1, 1 << 20, None)
    assert rc == -1 and "16384" in D._lib.last_error()
    with pytest.raises(D._lib.DctaError):
        D._lib.call("dcta_lfq_quantize", 1, 1, 1, 4, 1, 63, 1.0, None)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dct_autoencoder_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "dcta_oracle" not in src and "torch_dct_standin" not in src and "ref_shim" not in src, f


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_cuda(D):
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    with pytest.raises(D._lib.DctaError):
        fe.preprocess(torch.rand(3, 64, 64))
    with pytest.raises(D._lib.DctaError):
        D.util.rgb_to_ipt(torch.rand(3, 8, 8))
    with pytest.raises(D._lib.DctaError):
        D.LFQ(codebook_size=16, num_codebooks=2)(torch.rand(1, 4, 8), torch.ones(1, 4, dtype=torch.bool))


def test_constants_match_reference_fixture(D, golden):
    g = golden("constants")
    assert np.array_equal(D.util.Trgb2lms.numpy(), g["Trgb2lms"])
    assert np.array_equal(D.util.Tlms2rgb.numpy(), g["Tlms2rgb"])
    assert np.array_equal(D.util.Mipt.numpy(), g["Mipt"])
    assert np.array_equal(D.util.Mipt.inverse().numpy(), g["MiptInv"])
    np.testing.assert_allclose(D.util._basis_host(45, 20), O.dct_basis(45, 20), atol=1e-7)


def test_geometry_and_next_fit_match_oracle(D):
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    ofe = O.FeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    for h, w in [(512, 512), (1024, 1024), (256, 256), (300, 451), (14, 14), (27, 500)]:
        assert fe._get_crop_dims(h, w) == ofe._get_crop_dims(h, w)
    assert fe._geometry(512, 512) == (36, 36, 32, 32)
    assert fe._geometry(1024, 1024) == (73, 73, 32, 32)
    assert fe._geometry(256, 256) == (18, 18, 18, 18)
    with pytest.raises(AssertionError):
        fe._get_crop_dims(13, 100)
    rng = np.random.default_rng(0)
    fe.max_seq_len = ofe.max_seq_len = 1000
    ks = rng.integers(1, 1001, 200).tolist()
    st = fe._next_fit(ks)
    ost = ofe.group_by_max_seq_len(ks)
    assert st.rows == ost["groups"] and st.row == ost["group"] and st.seq_len == ost["seq_len"]
    with pytest.raises(AssertionError):
        fe._next_fit([1001])
    for beta in (0.0, 0.02, 0.004, 0.5):
        assert D.get_max_seq_length(32, 32, 3, beta) == O.get_max_seq_length(32, 32, 3, beta)
    assert [D.util.power_of_two(i) for i in range(0, 20)] == [O.power_of_two(i) for i in range(0, 20)]


def test_k_follows_the_reference_random_stream(D):
    import random
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.01, 32, 32, 512)
    ofe = O.FeatureExtractor(3, 14, 0.01, 32, 32, 512)
    random.seed(42)
    a = [fe._choose_k(3072) for _ in range(50)]
    random.seed(42)
    b = [ofe._choose_k(3072) for _ in range(50)]
    assert a == b and min(a) >= 1 and max(a) <= 512


def test_dct_patches_container(D):
    ids = torch.tensor([[0, 0, 1, 1, 0], [0, 1, 2, 0, 0]])
    pad = torch.tensor([[False, False, False, False, True], [False, False, False, True, True]])
    dp = D.DCTPatches(patches=torch.zeros(2, 5, 4), key_pad_mask=pad, batched_image_ids=ids,
                      patch_channels=torch.zeros(2, 5, dtype=torch.long),
                      patch_positions=torch.arange(20).reshape(2, 5, 2),
                      patch_sizes=[(1, 1)] * 5, original_sizes=[(4, 4)] * 5)
    op = O.Patches(dp.patches.numpy(), pad.numpy(), ids.numpy(), dp.patch_channels.numpy(),
                   dp.patch_positions.numpy(), [], [])
    assert dp._attn_mask is None                               # lazy
    assert np.array_equal(dp.attn_mask.numpy(), op.attn_mask)  # reference polarity, FE:580-584
    assert tuple(dp.attn_mask.shape) == (2, 1, 5, 5)
    assert dp.row_num_images() == [2, 3]
    assert torch.equal(dp.h_indices, dp.patch_positions[..., 0])
    cp = dp.shallow_copy()
    cp.patches = torch.ones(2, 5, 4)
    assert float(dp.patches.sum()) == 0.0 and cp.key_pad_mask is dp.key_pad_mask
    assert dp.to(torch.device("cpu")) is dp
    codes = torch.arange(2 * 5 * 3).reshape(2, 5, 3)
    objs = D.to_dict(dp, codes)
    assert len(objs) == 5 and [len(o["codes"]) for o in objs] == [2, 2, 1, 1, 1]
    assert objs[1]["codes"][0] == {"c": 0, "h": 4, "w": 5, "data": [6, 7, 8]}
    back, bc = D.from_dict(objs[0])
    assert bc.tolist() == [[0, 1, 2], [3, 4, 5]] and back.patch_positions.tolist() == [[[0, 1], [2, 3]]]


def test_segment_table_layout(D):
    fe = D.DCTAutoencoderFeatureExtractor(3, 4, 0.0, 4, 4, 48)
    from dct_autoencoder_b200.feature_extraction_dct_autoencoder import _SEG_DTYPE
    import ctypes
    assert ctypes.sizeof(D._lib.Segment) == _SEG_DTYPE.itemsize == 24
    assert [f[0] for f in D._lib.Segment._fields_] == list(_SEG_DTYPE.names)


# --------------------------------------------------------------------------- multi-rank protocol
_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "oracle"))
import dcta_oracle as O
from dct_autoencoder_b200.patchnorm import all_reduce_sum_, stats_sync_enabled
rank = int(os.environ["RANK"])
dist.init_process_group("gloo", rank=rank, world_size=2)
assert stats_sync_enabled()
C, H, W, p = 2, 3, 3, 2
z = p * p
n_pos = C * H * W
def shard(r):
    rng = np.random.default_rng(100 + r)
    T = 60 + 10 * r
    return (rng.standard_normal((T, z)).astype(np.float32), rng.integers(0, C, T), rng.integers(0, H, T), rng.integers(0, W, T))
x, c, h, w = shard(rank)
flat = c * H * W + h * W + w
# phase 1: what dcta_patchnorm_batch_median writes -- [batch_n | batch_median * batch_n]
packed = np.zeros(n_pos + n_pos * z, np.float32)
for pid in range(n_pos):
    rows = x[flat == pid]
    packed[pid] = len(rows)
    if len(rows):
        packed[n_pos + pid * z: n_pos + (pid + 1) * z] = np.partition(rows, (len(rows) - 1) // 2, axis=0)[(len(rows) - 1) // 2] * np.float32(len(rows))
t = torch.from_numpy(packed)
all_reduce_sum_(t)
g = t.numpy()
n0 = np.zeros(n_pos, np.float32)
median = (np.zeros((n_pos, z), np.float32) * n0[:, None] + g[n_pos:].reshape(n_pos, z)) / np.maximum(n0 + g[:n_pos], 1)[:, None]
# phase 2: sum |x - median|
ad = np.zeros((n_pos, z), np.float32)
np.add.at(ad, flat, np.abs(x - median[flat]))
t2 = torch.from_numpy(ad)
all_reduce_sum_(t2)
b = (np.ones((n_pos, z), np.float32) * n0[:, None] + (t2.numpy() / np.maximum(g[:n_pos], 1)[:, None]) * g[:n_pos, None]) / np.maximum(n0 + g[:n_pos], 1)[:, None]
# oracle: the reference rule applied sequentially to each rank's shard in rank order (SURVEY 8e)
pn = O.PatchNorm(H, W, p, C)
for r in range(2):
    xs, cs, hs, ws = shard(r)
    pn.update_stats(xs, cs, hs, ws)
assert np.array_equal(pn.n.reshape(-1), g[:n_pos])
np.testing.assert_allclose(median, pn.median.reshape(n_pos, z), rtol=1e-5, atol=1e-6)
# b: the sequential reference measures deviations around INTERMEDIATE medians, so it is compared with
# the rule itself evaluated on the concatenated shards: sum_all |x - median_global| / n_total
xa = np.concatenate([shard(r)[0] for r in range(2)])
fa = np.concatenate([shard(r)[1] * H * W + shard(r)[2] * W + shard(r)[3] for r in range(2)])
eb = np.zeros((n_pos, z), np.float64)
np.add.at(eb, fa, np.abs(xa - median[fa]).astype(np.float64))
eb = eb / np.maximum(g[:n_pos], 1)[:, None]
seen = g[:n_pos] > 0
np.testing.assert_allclose(b[seen], eb[seen], rtol=1e-5, atol=1e-6)
assert np.all(b[~seen] == 0)          # unseen positions: b goes 1 -> 0 on the first update (PN:146-148)
# and equals the single-process result on the concatenated batch when medians agree: check determinism
gathered = [torch.zeros_like(t2) for _ in range(2)]
dist.all_gather(gathered, torch.from_numpy(b.astype(np.float32)))
assert torch.equal(gathered[0], gathered[1])      # every rank ends with identical tables
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_patchnorm_stat_sync_protocol_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o


# ---------------------------------------------------------------- shard format
def _sample(i, k=5, z=4, dtype=torch.float16):
    g = torch.Generator().manual_seed(i)
    return {"__key__": f"{i:08}", "patches.pth": torch.randn(k, z, generator=g).to(dtype),
            "positions.pth": torch.randint(0, 9, (k, 2), generator=g), "channels.pth": torch.randint(0, 3, (k,), generator=g),
            "original_size.pyd": (17 + i, 23), "patch_size.pyd": (4, 5)}


def test_shard_writer_emits_the_webdataset_layout(D, tmp_path):
    """preproc_dataset.py:66-84: gzip tar, members '{key}.{name}.{ext}' adjacent per sample, .pth = torch.save,
    .pyd = pickle -- checked with nothing but tarfile / torch.load / pickle."""
    import io, pickle, tarfile
    with D.shards.ShardWriter(str(tmp_path / "%06d.tar"), maxsize=1e9, compress=True) as w:
        for i in range(3):
            w.write(_sample(i))
    assert [os.path.basename(p) for p in w.paths] == ["000000.tar"]
    assert open(w.paths[0], "rb").read(2) == b"\x1f\x8b"            # compress=True gzips whatever the name says
    with tarfile.open(w.paths[0], "r:gz") as tar:
        names = tar.getnames()
        assert names == [f"{i:08}.{n}" for i in range(3) for n in
                         ("patches.pth", "positions.pth", "channels.pth", "original_size.pyd", "patch_size.pyd")]
        pt = torch.load(io.BytesIO(tar.extractfile("00000001.patches.pth").read()))
        assert torch.equal(pt, _sample(1)["patches.pth"]) and pt.dtype == torch.float16
        assert pickle.loads(tar.extractfile("00000002.original_size.pyd").read()) == (19, 23)


def test_shard_rollover_reader_and_collate(D, tmp_path):
    one = D.shards.ShardWriter(str(tmp_path / "x%06d.tar")).write(_sample(0))
    with D.shards.ShardWriter(str(tmp_path / "%06d.tar"), maxsize=2.5 * one, maxcount=100) as w:
        for i in range(7):
            w.write(_sample(i))
    assert len(w.paths) == 3 and w.total == 7                       # a shard closes once it holds >= maxsize bytes
    for url in (str(tmp_path / "{000000..000002}.tar"), w.paths, str(tmp_path / "0*.tar")):
        rows = list(D.shards.load_preprocessed_dataset(url))
        assert len(rows) == 7
        for i, r in enumerate(rows):
            s = _sample(i)
            assert set(r) == {"patches", "positions", "channels", "original_sizes", "patch_sizes"}
            assert torch.equal(r["patches"], s["patches.pth"]) and torch.equal(r["positions"], s["positions.pth"])
            assert torch.equal(r["channels"], s["channels.pth"])
            assert r["original_sizes"] == s["original_size.pyd"] and r["patch_sizes"] == s["patch_size.pyd"]
    cols = list(D.shards.batched(D.shards.load_preprocessed_dataset(w.paths), 3))
    assert [len(c["patches"]) for c in cols] == [3, 3, 1]
    assert cols[1]["original_sizes"] == [(20, 23), (21, 23), (22, 23)]
    with D.shards.ShardWriter(str(tmp_path / "c%06d.tar"), maxcount=2) as w2:
        for i in range(5):
            w2.write(_sample(i))
    assert len(w2.paths) == 3
    # a sample that lost a member is skipped, like the reference's warn_and_continue handler
    with D.shards.ShardWriter(str(tmp_path / "p%06d.tar")) as w3:
        w3.write(_sample(0))
        bad = _sample(1)
        del bad["channels.pth"]
        w3.write(bad)
        w3.write(_sample(2))
    assert len(list(D.shards.load_preprocessed_dataset(w3.paths))) == 2


def test_fast_pth_writer_is_a_valid_torch_archive(D):
    """The shard writer's own .pth encoder: readable by torch.load(weights_only=True) and by zipfile (CRCs), stores
    exactly the slice it was given, and falls back to torch.save where it does not apply."""
    import io, zipfile
    big = torch.arange(6 * 40, dtype=torch.float32).reshape(6, 40)
    for dt in (torch.float16, torch.float32, torch.bfloat16, torch.int64, torch.uint8, torch.bool):
        x = big.to(dt)[2:5]                                   # a slice of a larger buffer, as in preprocess_to_shards
        blob = D.shards._save_tensor(x)
        assert D.shards._FAST_SAVE_OK[dt]
        y = torch.load(io.BytesIO(blob), weights_only=True)
        assert y.dtype == dt and torch.equal(y, x)
        assert y.untyped_storage().nbytes() == x.numel() * x.element_size()      # not the whole parent buffer
        z = zipfile.ZipFile(io.BytesIO(blob))
        assert z.testzip() is None and "archive/data.pkl" in z.namelist() and "archive/data/0" in z.namelist()
    for odd in (big.t(), torch.empty(0, 3)):                  # non-contiguous / empty: plain torch.save
        y = torch.load(io.BytesIO(D.shards._save_tensor(odd)), weights_only=True)
        assert y.shape == odd.shape and torch.equal(y, odd)


def test_numa_binding_is_a_noop_without_a_gpu(D):
    """util.bind_to_gpu_numa must never raise or change the affinity when the topology cannot be read."""
    if torch.cuda.is_available():
        pytest.skip("needs a host without CUDA")
    before = os.sched_getaffinity(0)
    assert D.util.gpu_numa_cpus() is None
    assert D.util.bind_to_gpu_numa() is False
    assert os.sched_getaffinity(0) == before


def test_to_dict_from_dict_match_the_reference_fixture(D, golden):
    """to_dict / from_dict (DP:54-122) against the UNMODIFIED reference's output on a packed 3-image batch
    (tests/golden/make_golden_configs.py).  Host tensors: these two functions are plain serialisation.  Also covers
    the code-only batches of the fused encode path (patches=None), which `to`, `repr` and `to_dict` must accept."""
    import json
    g = golden("to_dict")
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "to_dict.json")))
    t = torch.from_numpy
    for patches in (torch.zeros(tuple(g["patches_shape"])), None):
        dp = D.DCTPatches(patches=patches, key_pad_mask=t(g["key_pad_mask"]), batched_image_ids=t(g["image_ids"]),
                          patch_channels=t(g["channels"]), patch_positions=t(g["positions"]),
                          patch_sizes=[tuple(x) for x in g["patch_sizes"].tolist()],
                          original_sizes=[tuple(x) for x in g["original_sizes"].tolist()])
        got = D.to_dict(dp, t(g["codes"]))
        assert json.loads(json.dumps(got)) == want
        assert "DCTPatches(" in repr(dp)
        assert dp.to("cpu") is dp
    dp1, codes1 = D.from_dict(want[1])
    assert torch.equal(codes1, t(g["fd_codes"])) and codes1.dtype == torch.int64
    for ours, key in ((dp1.patch_channels, "fd_channels"), (dp1.patch_positions, "fd_positions"),
                      (dp1.key_pad_mask, "fd_key_pad_mask"), (dp1.batched_image_ids, "fd_image_ids"),
                      (dp1.attn_mask, "fd_attn_mask"), (dp1.patches, "fd_patches")):
        assert ours.dtype == t(g[key]).dtype and torch.equal(ours, t(g[key])), key
    assert dp1.patch_sizes == [want[1]["size"]] and dp1.original_sizes == [want[1]["original_size"]]
    # round trip through our own pair
    assert D.to_dict(dp1, codes1[None])[0]["codes"] == want[1]["codes"]


def test_caller_supplied_token_counts_are_validated(D):
    """process_batch / process_batch_to_codes index order[img, :k]: k beyond the image's own token count (or a wrong
    number of ks) must be refused on the host, before any kernel sees it."""
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    n_tok = 9 * 8 * 3
    assert fe._check_ks([1, n_tok], 2, n_tok) == [1, n_tok]
    for bad in ([n_tok + 1, 5], [0, 5], [5], [5, 5, 5]):
        with pytest.raises(AssertionError):
            fe._check_ks(bad, 2, n_tok)
    with pytest.raises(AssertionError):
        D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 100)._check_ks([101], 1, n_tok)


def test_table_cache_is_keyed_by_device_and_evicts_lru(D):
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    fe._table_cache_size = 3
    made = []
    orig = torch.Tensor.pin_memory
    # no GPU here: exercise the cache logic with CPU "device" copies
    torch.Tensor.pin_memory = lambda self: self
    try:
        for i in range(5):
            fe._cached_table(("cpu", bytes([i])), np.full(4, i, np.uint8), "cpu")
        assert list(k[1] for k in fe._table_cache) == [bytes([2]), bytes([3]), bytes([4])]
        a = fe._cached_table(("cpu", bytes([2])), np.full(4, 2, np.uint8), "cpu")      # hit: becomes most recent
        fe._cached_table(("cpu", bytes([9])), np.full(4, 9, np.uint8), "cpu")
        assert bytes([2]) in [k[1] for k in fe._table_cache] and bytes([3]) not in [k[1] for k in fe._table_cache]
        b = fe._cached_table(("other", bytes([2])), np.full(4, 2, np.uint8), "cpu")     # same bytes, other device
        assert a is not b
        fe._keepalive = []
        c = fe._cached_table(("cpu", bytes([2])), np.full(4, 2, np.uint8), "cpu")
        assert fe._keepalive == [c] and c is a
    finally:
        torch.Tensor.pin_memory = orig


@pytest.mark.parametrize("n,k", [(512, 448), (256, 252), (1024, 448), (90, 84), (16, 16), (451, 448), (301, 294), (15, 14)])
def test_basis_init_matches_the_float64_definition(D, n, k):
    """dcta_basis_init (host arithmetic in libdcta, no GPU): every layout equals the orthonormal DCT-II definition
    evaluated with numpy in float64 -- fp32 table to the last bit, hi + lo of the split tables to 2^-21 relative
    (and hi/lo themselves bit for bit wherever numpy's cos and libm's agree to the double's last place)."""
    lib = D._lib.load()
    L = D._lib
    q = np.arange(k, dtype=np.float64)[:, None]
    m = np.arange(n, dtype=np.float64)[None, :]
    c = np.cos(np.pi * (2 * m + 1) * q / (2 * n)) * np.sqrt(2.0 / n)
    c[0, :] = np.sqrt(1.0 / n)
    from dct_autoencoder_b200.util import _basis_tables, _round8
    f32, _, _ = _basis_tables(L.BASIS_F32, n, k, (k, n), False)
    assert np.abs(f32.astype(np.float64) - c).max() <= 2.0 ** -24 * np.abs(c).max() * 1.01
    assert (f32 != c.astype(np.float32)).mean() < 1e-3
    hi, lo, rs = _basis_tables(L.BASIS_SPLIT_FWD, n, k, (k, _round8(n)), True)
    got = (hi.astype(np.float64) + lo.astype(np.float64))[:, :n] * rs.astype(np.float64)[:, None]
    assert np.abs(got - c).max() <= 2.0 ** -21 * np.abs(c).max()
    assert np.all(hi[0, :n] == 32.0) and np.all(lo[0, :n] == 0.0) and not hi[:, n:].any()
    hi_t, lo_t, _ = _basis_tables(L.BASIS_SPLIT_INV, n, k, (n, _round8(k)), False)
    got = (hi_t.astype(np.float64) + lo_t.astype(np.float64))[:, :k].T / 1024.0
    assert np.abs(got - c).max() <= 2.0 ** -21 * np.abs(c).max()
    if k % 2 == 0:
        # folded layouts: ceil(n/2) samples per group, rows padded to 8; odd n: the middle sample pairs with itself and
        # the odd rows hold an exact 0 there (cos(pi q / 2))
        n2, ld = (n + 1) // 2, _round8((n + 1) // 2)
        fh, fl, frs = _basis_tables(L.BASIS_FOLD_FWD, n, k, (2, k // 2, ld), True)
        got = (fh.astype(np.float64) + fl.astype(np.float64))[:, :, :n2] * frs.reshape(2, k // 2, 1).astype(np.float64)
        want = np.stack([c[0::2, :n2], c[1::2, :n2]])
        assert np.abs(got - want).max() <= 2.0 ** -21 * np.abs(c).max()
        assert not fh[:, :, n2:].any() and not fl[:, :, n2:].any()
        if n % 2:
            assert not fh[1, :, n2 - 1].any() and not fl[1, :, n2 - 1].any()
        ih, il, _ = _basis_tables(L.BASIS_FOLD_INV, n, k, (2, n2, _round8(k // 2)), False)
        got = np.transpose((ih.astype(np.float64) + il.astype(np.float64))[:, :, :k // 2], (0, 2, 1)) / 1024.0
        assert np.abs(got - want).max() <= 2.0 ** -21 * np.abs(c).max()
    assert lib.dcta_basis_elems(99, n, k) == -1
    assert lib.dcta_basis_init(L.BASIS_F32, 8, 9, f32.ctypes.data, None, None) == -1      # k > n


def test_loader_host_logic(tmp_path):
    """dataset.py host side (no GPU): torchvision's smaller-edge rule, shard URL expansion, webdataset sample grouping."""
    import io
    import json
    import tarfile
    from torchvision import transforms
    from dct_autoencoder_b200 import dataset as DS
    for h, w, size in [(900, 1300, 531), (1300, 900, 531), (1000, 1000, 768), (769, 40, 39), (333, 1001, 255)]:
        want = tuple(transforms.Resize(size)(torch.zeros(1, h, w)).shape[-2:])
        assert DS._resize_smaller_edge(h, w, size) == want, (h, w, size)
    assert DS.expand_urls("a/s-{0008..0011}.tar") == ["a/s-0008.tar", "a/s-0009.tar", "a/s-0010.tar", "a/s-0011.tar"]
    path = os.path.join(tmp_path, "x-0000.tar")
    with tarfile.open(path, "w") as tf:
        for key, members in (("dir/aaa", {"jpg": b"J1", "json": json.dumps({"height": 5, "width": 6}).encode(), "txt": b"t"}),
                             ("dir/bbb", {"jpeg": b"J2", "json": b"{}"})):
            for ext, data in members.items():
                ti = tarfile.TarInfo(f"{key}.{ext}")
                ti.size = len(data)
                tf.addfile(ti, io.BytesIO(data))
    got = list(DS.iter_tar_samples(os.path.join(tmp_path, "x-*.tar")))
    assert [g["__key__"] for g in got] == ["dir/aaa", "dir/bbb"]
    assert got[0]["jpg"] == b"J1" and json.loads(got[0]["json"])["width"] == 6 and "txt" not in got[0]
    assert got[1]["jpeg"] == b"J2"
    assert DS.dict_collate([{"a": 1, "b": 2}, {"a": 3, "b": 4}]) == {"a": [1, 3], "b": [2, 4]}
    assert DS.tuple_collate([(1, 2), (3, 4)]) == [[1, 3], [2, 4]]
    fe_like = type("P", (), dict(patch_size=14, max_patch_w=32, max_patch_h=32))()
    assert DS.max_image_size(fe_like) == 768
    with pytest.raises(Exception):
        DS.decode_jpegs([b"x"], "cpu")          # the loader delivers to the GPU: no CPU product path


def test_extractor_save_and_load_preprocessor_config(D, tmp_path):
    """FE:80: the reference inherits FeatureExtractionMixin; the same preprocessor_config.json round-trips here, and
    transformers' own mixin reads the file we write (same keys, "feature_extractor_type")."""
    import json
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 2.5, 32, 24, 1024, channel_importances=(4.0, 1.0, 2.0),
                                          patch_sample_magnitude_weight=0.25)
    files = fe.save_pretrained(tmp_path / "proc")
    assert [os.path.basename(f) for f in files] == ["preprocessor_config.json"]
    cfg = json.load(open(files[0]))
    assert cfg["feature_extractor_type"] == "DCTAutoencoderFeatureExtractor" and cfg["max_patch_w"] == 24
    for src in (tmp_path / "proc", files[0]):
        fe2 = D.DCTAutoencoderFeatureExtractor.from_pretrained(src)
        assert fe2.to_dict() == fe.to_dict()
        assert torch.equal(fe2.channel_importances, fe.channel_importances)
    fe3 = D.DCTAutoencoderFeatureExtractor.from_pretrained(tmp_path / "proc", max_seq_len=77)
    assert fe3.max_seq_len == 77 and fe3.patch_size == 14
    with pytest.raises(EnvironmentError):
        D.DCTAutoencoderFeatureExtractor.from_pretrained(tmp_path / "nothing-here")
    with pytest.raises(ValueError):
        D.DCTAutoencoderFeatureExtractor.from_dict({"channels": 3})
    # the file is what transformers' mixin expects
    from transformers.feature_extraction_utils import FeatureExtractionMixin
    d, _ = FeatureExtractionMixin.get_feature_extractor_dict(str(tmp_path / "proc"))
    assert d["patch_size"] == 14 and d["channel_importances"] == [4.0, 1.0, 2.0]
