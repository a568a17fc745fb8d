"""acmil_b200/h5bag.py: the reference's feature-bag container (Step2_feature_extract.py:164-167 writer,
datasets/datasets.py:16-43, 138-155 reader) without h5py -- round trips, the on-disk structure the HDF5 specification
prescribes for the default ("earliest") format, the reference's loader on top of it, and (when h5py happens to be
installed) a cross-check of writer and reader against libhdf5 itself."""
import struct

import numpy as np
import pytest

from acmil_b200 import h5bag as H


def _bags(n_slides, seed=0, dim=384):
    rng = np.random.default_rng(seed)
    out = {}
    for i in range(n_slides):
        n = int(rng.integers(1, 40))
        name = ("test_%03d" if i % 3 == 0 else "tumor_%03d") % i
        out[name] = (rng.standard_normal((n, dim)).astype(np.float32), rng.integers(0, 100000, (n, 2)).astype(np.int64), np.int64(i % 2))
    return out


@pytest.mark.parametrize("n_slides", [1, 7, 9, 70, 400])
def test_round_trip(tmp_path, n_slides):
    bags = _bags(n_slides, seed=n_slides)
    path = tmp_path / "bags.h5"
    H.write_bags(path, bags)
    with H.H5BagFile(path) as f:
        assert sorted(f.keys()) == sorted(bags)
        for name, (feat, coords, label) in bags.items():
            g = f[name]
            assert sorted(g.keys()) == ["coords", "feat"]
            got = g["feat"][:]
            assert got.dtype == np.float16 and got.shape == feat.shape      # stored as fp16 (Step2_feature_extract.py:165)
            np.testing.assert_array_equal(got, feat.astype(np.float16))
            np.testing.assert_array_equal(g["coords"][:], coords)
            assert g["coords"][:].dtype == np.int64
            assert int(g.attrs["label"]) == int(label)
            np.testing.assert_array_equal(g["feat"][2:5], feat.astype(np.float16)[2:5])


def test_on_disk_structure_follows_the_specification(tmp_path):
    """Superblock version 0 at offset 0; the root symbol-table entry points at a version-1 object header whose symbol-table
    message names a 'TREE' node and a 'HEAP'; names are null-terminated in the heap; a dataset header carries dataspace,
    datatype (IEEE half: 16 bits, exponent at bit 10 with 5 bits and bias 15), fill value and a contiguous layout."""
    path = tmp_path / "one.h5"
    feat = np.arange(12, dtype=np.float32).reshape(3, 4)
    H.write_bags(path, {"slide_a": (feat, np.zeros((3, 2), np.int32), 1)})
    b = open(path, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0 and b[13] == 8 and b[14] == 8
    leaf_k, internal_k = struct.unpack_from("<HH", b, 16)
    assert leaf_k >= 1 and internal_k == 16
    base, free, eof, drv = struct.unpack_from("<QQQQ", b, 24)
    assert base == 0 and free == H.UNDEF and eof == len(b) and drv == H.UNDEF
    name_off, root_hdr, cache, _r, btree, heap = struct.unpack_from("<QQIIQQ", b, 56)
    assert name_off == 0 and cache == 1 and b[btree:btree + 4] == b"TREE" and b[heap:heap + 4] == b"HEAP"
    ver, _r, nmsg, refs, size = struct.unpack_from("<BBHII", b, root_hdr)
    assert ver == 1 and nmsg == 1 and refs == 1 and size % 8 == 0
    mtype, msize = struct.unpack_from("<HH", b, root_hdr + 16)
    assert mtype == 0x0011 and struct.unpack_from("<QQ", b, root_hdr + 24) == (btree, heap)
    _ds, _fr, heap_data = struct.unpack_from("<QQQ", b, heap + 8)
    assert b[heap_data + 8:heap_data + 16] == b"slide_a\0"
    # every structure starts on an 8-byte boundary
    assert all(a % 8 == 0 for a in (root_hdr, btree, heap, heap_data))
    dt = H._dtype_message(np.float16)
    assert dt[0] == 0x11 and dt[1] == 0x20 and dt[2] == 15 and struct.unpack_from("<I", dt, 4)[0] == 2
    assert struct.unpack_from("<HHBBBBI", dt, 8) == (0, 16, 10, 5, 0, 10, 15)
    assert struct.unpack_from("<HHBBBBI", H._dtype_message(np.float32), 8) == (0, 32, 23, 8, 0, 23, 127)
    di = H._dtype_message(np.int64)
    assert di[0] == 0x10 and di[1] == 0x08 and struct.unpack_from("<IHH", di, 4) == (8, 0, 64)


def test_reader_rejects_what_it_does_not_implement(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not hdf5" * 20)
    with pytest.raises(ValueError):
        H.H5BagFile(p)
    good = tmp_path / "g.h5"
    H.write_bags(good, _bags(2))
    raw = bytearray(good.read_bytes())
    raw[8] = 2                                             # pretend: superblock version 2 (libver='latest')
    bad = tmp_path / "v2.h5"
    bad.write_bytes(bytes(raw))
    with pytest.raises(NotImplementedError):
        H.H5BagFile(bad)
    with pytest.raises(ValueError):
        H.H5BagFile(good, "w")


def test_reference_loader_on_top(tmp_path):
    """split_dataset_camelyon / HDF5_feat_dataset2 (datasets/datasets.py:16-43, 138-155)."""
    from acmil_b200 import Struct
    bags = _bags(12, seed=3, dim=16)
    path = tmp_path / "c16.h5"
    H.write_bags(path, bags)
    tr, tr_names, va, va_names, te, te_names = H.split_dataset_camelyon(path, Struct(seed=1, dataset="camelyon"))
    assert set(te_names) == {n for n in bags if "test" in n} and len(va_names) >= 1
    assert set(tr_names) | set(va_names) | set(te_names) == set(bags) and not (set(tr_names) & set(va_names))
    ds = H.HDF5_feat_dataset2(tr, tr_names)
    item = ds[0]
    assert set(item) == {"input", "coords", "label"} and item["input"].dtype == np.float16
    np.testing.assert_array_equal(item["input"], bags[tr_names[0]][0].astype(np.float16))
    split = {"train_names": list(bags)[:3], "val_names": list(bags)[3:5], "test_names": list(bags)[5:]}
    tr2, n2, *_ = H.split_dataset_camelyon(path, Struct(seed=1), split)
    assert n2 == split["train_names"] and set(tr2) == set(n2)


def test_cross_check_against_h5py(tmp_path):
    """Only where h5py exists (not in the build image): libhdf5 reads what write_bags wrote, and H5BagFile reads what the
    reference's writer code produces with h5py."""
    h5py = pytest.importorskip("h5py")
    bags = _bags(20, seed=9, dim=32)
    mine = tmp_path / "mine.h5"
    H.write_bags(mine, bags)
    with h5py.File(mine, "r") as f:
        assert sorted(f.keys()) == sorted(bags)
        for name, (feat, coords, label) in bags.items():
            np.testing.assert_array_equal(f[name]["feat"][:], feat.astype(np.float16))
            np.testing.assert_array_equal(f[name]["coords"][:], coords)
            assert int(f[name].attrs["label"]) == int(label)
    theirs = tmp_path / "theirs.h5"
    with h5py.File(theirs, "w") as f:                       # Step2_feature_extract.py:163-167
        for name, (feat, coords, label) in bags.items():
            g = f.create_group(name)
            g.create_dataset("feat", data=feat.astype(np.float16))
            g.create_dataset("coords", data=coords)
            g.attrs["label"] = label
    with H.H5BagFile(theirs) as f:
        for name, (feat, coords, label) in bags.items():
            np.testing.assert_array_equal(f[name]["feat"][:], feat.astype(np.float16))
            np.testing.assert_array_equal(f[name]["coords"][:], coords)
            assert int(f[name].attrs["label"]) == int(label)


@pytest.mark.gpu
def test_prefetcher_feeds_the_head(tmp_path):
    """H5 file -> loader -> BagPrefetcher (pinned double buffering, fp16 on the device) -> ACMIL_GA.forward_bags: same logits as
    the widened fp32 bag through the module's forward."""
    import torch
    from acmil_b200 import ACMIL_GA, Struct
    bags = _bags(5, seed=4, dim=384)
    path = tmp_path / "b.h5"
    H.write_bags(path, bags)
    tr, names, *_ = H.split_dataset_camelyon(path, Struct(seed=0), {"train_names": sorted(bags), "val_names": [], "test_names": []})
    ds = H.HDF5_feat_dataset2(tr, names)
    torch.manual_seed(2)
    m = ACMIL_GA(Struct(D_feat=384, D_inner=128, n_class=2, n_token=5), n_token=5, n_masked_patch=10, mask_drop=0.6).cuda().eval()
    seen = []
    with torch.no_grad():
        for name, x16, coords, label in H.BagPrefetcher(ds, "cuda"):
            assert x16.dtype == torch.float16 and x16.is_cuda
            _, slide, _ = m.forward_bags(x16, [0, x16.shape[0]])
            ref = m(torch.from_numpy(bags[name][0].astype(np.float16)).float().cuda()[None])[1]
            np.testing.assert_allclose(slide.cpu().numpy(), ref.cpu().numpy(), rtol=1e-5, atol=1e-6)
            seen.append(name)
    assert seen == names
