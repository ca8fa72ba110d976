"""GPU tests of the offline shard path: batched preprocess -> shards -> iter_batches."""
import random

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def D():
    import dct_autoencoder_b200 as d
    d._lib.load()
    return d


@pytest.mark.parametrize("size,patch,beta", [((96, 128), 8, 0.0), ((140, 126), 14, 0.05), ((64, 64), 16, 0.0)])
def test_preprocess_batch_equals_per_image_preprocess(D, size, patch, beta):
    torch.manual_seed(3)
    x = torch.rand(5, 3, *size).cuda()
    fe = D.DCTAutoencoderFeatureExtractor(3, patch, beta, 8, 8, 120)
    random.seed(11)
    one = [fe.preprocess(im) for im in x]
    random.seed(11)                                   # the k draws consume Python's global stream in image order
    many = fe.preprocess_batch(x)
    assert len(many) == len(one)
    for a, b in zip(one, many):
        for key in ("patches", "positions", "channels"):
            assert torch.equal(a[key], b[key]), key
        assert a["original_sizes"] == b["original_sizes"] and a["patch_sizes"] == b["patch_sizes"]
    if beta > 0:
        assert len({r["patches"].shape[0] for r in many}) > 1


def test_preprocess_batch_keeps_the_input_dtype(D):
    x = torch.rand(2, 3, 64, 80).cuda()
    fe = D.DCTAutoencoderFeatureExtractor(3, 8, 0.0, 8, 8, 64)
    r16 = fe.preprocess_batch(x.half())
    r32 = fe.preprocess_batch(x.half().float())
    assert r16[0]["patches"].dtype == torch.float16
    assert torch.equal(r16[1]["patches"], r32[1]["patches"].half())


@pytest.mark.parametrize("dtype", [None, torch.float16])
def test_shards_feed_iter_batches_like_the_live_encoder(D, tmp_path, dtype):
    """write shards -> load_preprocessed_dataset -> dict_collate -> iter_batches == process_batch on the same images."""
    torch.manual_seed(5)
    fe = D.DCTAutoencoderFeatureExtractor(3, 8, 0.0, 8, 8, 200)
    batches = [torch.rand(4, 3, 72, 104), torch.rand(3, 3, 72, 104).cuda(), torch.rand(2, 3, 72, 104)]
    ks = [[50, 80, 64, 7], [192, 1, 100], [30, 170]]
    info = D.shards.preprocess_to_shards(batches, fe, str(tmp_path), dtype=dtype, maxsize=40e3, compress=True, ks=ks)
    assert info["samples"] == 9 and len(info["shards"]) > 1
    rows = list(D.shards.load_preprocessed_dataset(str(tmp_path)))
    assert [r["patches"].shape[0] for r in rows] == [k for kk in ks for k in kk]
    assert all(r["patches"].dtype == (dtype or torch.float32) and not r["patches"].is_cuda for r in rows)
    assert all(r["original_sizes"] == (72, 104) and r["patch_sizes"] == (9, 13) for r in rows)
    got = next(fe.iter_batches(D.shards.batched(iter(rows), 9), batch_size=None))   # batch_size=None is single-shot
    want = fe.process_batch(torch.cat([b.cuda() for b in batches]), [k for kk in ks for k in kk])
    wp = want.patches if dtype is None else want.patches.to(dtype)
    assert torch.equal(got.patches.to(wp.device), wp)
    for f in ("key_pad_mask", "batched_image_ids", "patch_channels", "patch_positions"):
        assert torch.equal(getattr(got, f).to(wp.device), getattr(want, f)), f
    assert got.original_sizes == want.original_sizes and got.patch_sizes == want.patch_sizes


def test_writer_failure_reaches_the_caller(D, tmp_path):
    fe = D.DCTAutoencoderFeatureExtractor(3, 8, 0.0, 8, 8, 64)
    target = tmp_path / "file"
    target.write_text("not a directory")
    with pytest.raises(Exception):
        D.shards.preprocess_to_shards([torch.rand(2, 3, 64, 64)], fe, str(target / "sub"))


def test_parallel_writers_store_every_sample_once(D, tmp_path):
    fe = D.DCTAutoencoderFeatureExtractor(3, 8, 0.0, 8, 8, 64)
    torch.manual_seed(2)
    batches = [torch.rand(3, 3, 64, 64) for _ in range(6)]
    info = D.shards.preprocess_to_shards(batches, fe, str(tmp_path), dtype=torch.float16, compress=True, writers=3,
                                         maxsize=60e3)
    assert info["samples"] == 18
    rows = {r["__key__"]: r for r in D.shards.iter_samples(info["shards"])}
    assert sorted(rows) == [f"{i:08}" for i in range(18)]
    want = fe.preprocess_batch(torch.cat(batches).cuda())
    for i, w in enumerate(want):
        got = rows[f"{i:08}"]
        assert torch.equal(got["patches.pth"], w["patches"].half().cpu())
        assert torch.equal(got["positions.pth"], w["positions"].cpu())
