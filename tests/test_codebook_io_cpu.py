"""On-disk codebook format (opengaussian_b200.codebook_io) against files written and read by the reference's own
save_kmeans / load_code_book (tests/golden/make_codebook_golden.py), plus round trips at awkward sizes."""
import importlib.util
import os
import tempfile
import types

import numpy as np
import pytest
import torch

from opengaussian_b200 import codebook_io as cio

GOLD = os.path.join(os.path.dirname(__file__), "golden", "codebook_golden.npz")


def _gm():
    spec = importlib.util.spec_from_file_location("mcg", os.path.join(os.path.dirname(GOLD), "make_codebook_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("name", ["root_1000", "leaf_777"])
def test_files_identical_to_reference(name):
    m, gold = _gm(), np.load(GOLD)
    ids, centers, mode = m.inputs(name)
    km = types.SimpleNamespace(cls_ids=ids, leaf_cls_ids=ids, centers=centers, leaf_centers=centers)
    with tempfile.TemporaryDirectory() as d:
        cio.save_kmeans([km], ["ins_feat"], d, mode=mode)
        sub = os.path.join(d, f"{mode}_code_book")
        assert sorted(os.listdir(sub)) == ["kmeans_args.npy", "kmeans_centers.pth", "kmeans_inds.bin"]
        assert np.array_equal(np.fromfile(os.path.join(sub, "kmeans_inds.bin"), dtype=np.uint8), gold[f"{name}/inds"])
        args = np.load(os.path.join(sub, "kmeans_args.npy"), allow_pickle=True).item()
        assert args == {"params": ["ins_feat"], "n_bits": int(gold[f"{name}/n_bits"]), "total_len": int(gold[f"{name}/total_len"])}
        codebook, loaded = cio.load_code_book(sub)
        assert np.array_equal(loaded, gold[f"{name}/loaded"]) and np.array_equal(loaded, ids.numpy())
        assert np.array_equal(codebook["ins_feat"].numpy(), gold[f"{name}/centers"])


def test_loader_reads_reference_bytes():
    """Files as the reference wrote them (golden bytes) decode to the reference loader's indices."""
    gold = np.load(GOLD)
    for name in ("root_1000", "leaf_777"):
        got = cio.unpack_ids(gold[f"{name}/inds"], int(gold[f"{name}/n_bits"]), int(gold[f"{name}/total_len"]))
        assert np.array_equal(got.numpy(), gold[f"{name}/loaded"])


@pytest.mark.parametrize("n,k", [(1, 1), (2, 2), (9, 641), (12345, 64), (1 << 20, 641), ((1 << 20) + 3, 10)])
def test_pack_round_trip(n, k):
    rs = np.random.RandomState(n % 1000)
    ids = torch.from_numpy(rs.randint(0, k, size=n).astype(np.int64))
    n_bits = max(int(np.ceil(np.log2(n))), 1) if n > 1 else 1
    if k - 1 >= (1 << n_bits):
        pytest.skip("ids do not fit the reference's ceil(log2(#points)) bits")
    raw = cio.pack_ids(ids, n_bits)
    assert raw.dtype == np.uint8 and raw.size == (n * n_bits + 7) // 8
    want = np.packbits(cio.dec2binary(ids, n_bits).numpy().reshape(-1).astype(np.uint8), bitorder="big")
    assert np.array_equal(raw, want)
    assert np.array_equal(cio.unpack_ids(raw, n_bits, n * n_bits).numpy(), ids.numpy())


def test_two_parameters_share_one_bit_string():
    """kmeans_list with two entries whose first part does not end on a byte boundary (reference concatenates bits)."""
    a = torch.arange(5, dtype=torch.int64) % 3          # 5 ids x 3 bits = 15 bits
    b = torch.arange(5, dtype=torch.int64) % 5
    kms = [types.SimpleNamespace(cls_ids=a, centers=torch.zeros(3, 6)), types.SimpleNamespace(cls_ids=b, centers=torch.zeros(5, 6))]
    with tempfile.TemporaryDirectory() as d:
        cio.save_kmeans(kms, ["ins_feat", "other"], d, mode="root")
        sub = os.path.join(d, "root_code_book")
        raw = np.fromfile(os.path.join(sub, "kmeans_inds.bin"), dtype=np.uint8)
        bits = np.concatenate([cio.dec2binary(a, 3).numpy().reshape(-1), cio.dec2binary(b, 3).numpy().reshape(-1)]).astype(np.uint8)
        assert np.array_equal(raw, np.packbits(bits, bitorder="big"))
        _, first = cio.load_code_book(sub)
        assert np.array_equal(first, a.numpy())


def test_cluster_lang_file(tmp_path):
    import numpy as np
    import torch
    from opengaussian_b200 import codebook_io as cio
    g = torch.Generator().manual_seed(3)
    feat = torch.randn(20, 512, generator=g)
    score = torch.rand(20, generator=g)
    occu = torch.randint(0, 12, (20,), generator=g).float()
    ind = torch.randint(0, 20, (1000,), generator=g)
    cio.save_cluster_lang(str(tmp_path), feat, score, occu, ind)
    raw = np.load(str(tmp_path / "cluster_lang.npz"))            # the reference's readers use these member names
    assert sorted(raw.files) == sorted(cio.CLUSTER_LANG_KEYS)
    assert np.array_equal(raw["leaf_feat.npy"], feat.numpy()) and raw["leaf_ind"].dtype == np.int64
    f2, s2, o2, i2 = cio.load_cluster_lang(str(tmp_path), device="cpu")
    assert torch.equal(i2, ind) and torch.equal(s2, score) and torch.equal(o2, occu)
    assert torch.equal(f2[occu >= 5], feat[occu >= 5]) and not f2[occu < 5].any()
