"""The reference's feature-bag container: one HDF5 file, one group per slide with

    feat    [N, D]  float16      Step2_feature_extract.py:165   slide_grp.create_dataset('feat', data=...astype(np.float16))
    coords  [N, 2]  integer      Step2_feature_extract.py:166
    attrs['label']  integer      Step2_feature_extract.py:167

read back by datasets/datasets.py:16-43 (``split_dataset_camelyon``: every slide's feat / coords / label into a dict) and
served by ``HDF5_feat_dataset2`` (datasets.py:138-155).  h5py is not part of this image, so this module carries its own
implementation of exactly the HDF5 subset that layout needs -- what h5py / libhdf5 write with default settings ("earliest"
format: version-0 superblock, version-1 object headers, symbol-table groups with a B-tree and a local heap, contiguous or
compact dataset storage, header-resident attributes) -- following the public HDF5 File Format Specification, version 1.1:

    write_bags(path, {slide: (feat, coords, label)})      the writer of Step2 (no h5py needed)
    H5BagFile(path)                                        read-only: keys(), f[slide]['feat'][:], f[slide].attrs['label']
    split_dataset_camelyon / HDF5_feat_dataset2            the reference's loader on top of it
    BagPrefetcher                                          pinned-memory double buffering: fp16 bags on the GPU, ready
                                                           for ACMIL_GA.forward_bags (the kernels read fp16 rows as they are)

When h5py IS importable, ``H5BagFile`` and ``store_features`` (extract.py) use it, and tests/test_h5bag.py cross-checks this
writer against it; without h5py the format parity is pinned only by round trips and by the structure checks of the test
("parity unpinned" at the libhdf5 boundary).  Not supported (raises): chunked / compressed datasets, the "latest" file
format (version-2 object headers, fractal-heap groups), more than one B-tree level per group (> 32 * 2 * leaf_k entries).
"""
from __future__ import annotations

import struct
from typing import Dict, Iterable, Optional, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
INTERNAL_K = 16                     # B-tree nodes of a group hold up to 2 K children


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


# ------------------------------------------------------------------------------------------ datatype / dataspace messages
def _dtype_message(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.byteorder == ">":
        raise ValueError("big-endian arrays are not supported")
    if dt.kind == "f":
        exp_size, mant = {2: (5, 10), 4: (8, 23), 8: (11, 52)}[dt.itemsize]
        bits = dt.itemsize * 8
        # class 1 (floating point), version 1; bit field: little-endian, mantissa normalisation 2 (msb implied), sign bit position
        head = struct.pack("<BBBBI", 0x11, 0x20, bits - 1, 0, dt.itemsize)
        return head + struct.pack("<HHBBBBI", 0, bits, mant, exp_size, 0, mant, (1 << (exp_size - 1)) - 1)
    if dt.kind in "iu":
        head = struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize)
        return head + struct.pack("<HH", 0, dt.itemsize * 8)
    raise ValueError(f"unsupported dtype {dt}")


def _parse_dtype(b: bytes) -> np.dtype:
    cls, bf0, _bf1, _bf2, size = struct.unpack_from("<BBBBI", b, 0)
    klass = cls & 0x0F
    if bf0 & 1:
        raise NotImplementedError("big-endian datasets are not supported")
    if klass == 1:
        return np.dtype({2: np.float16, 4: np.float32, 8: np.float64}[size])
    if klass == 0:
        return np.dtype(("i" if bf0 & 0x08 else "u") + str(size))
    raise NotImplementedError(f"HDF5 datatype class {klass} is not supported (feat / coords / label are numeric)")


def _dataspace_message(shape: Tuple[int, ...]) -> bytes:
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(s)) for s in shape)


def _parse_dataspace(b: bytes) -> Tuple[int, ...]:
    ver, rank, flags = struct.unpack_from("<BBB", b, 0)
    off = 8 if ver == 1 else 4
    return tuple(struct.unpack_from("<Q", b, off + 8 * i)[0] for i in range(rank))


def _message(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _object_header(messages: Iterable[bytes]) -> bytes:
    msgs = list(messages)
    body = b"".join(msgs)
    return struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body


def _attribute_message(name: str, value) -> bytes:
    arr = np.asarray(value)
    if arr.dtype.kind not in "iuf":
        raise ValueError("only numeric attributes are supported")
    nm = name.encode() + b"\0"
    dtm, dsm = _dtype_message(arr.dtype), _dataspace_message(arr.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dtm), len(dsm)) + _pad8(nm) + _pad8(dtm) + _pad8(dsm) + arr.tobytes()
    return _message(0x000C, body)


# ------------------------------------------------------------------------------------------ writer
class _Writer:
    def __init__(self, leaf_k: int):
        self.buf = bytearray(96)          # the superblock is written last
        self.leaf_k = leaf_k

    def tell(self) -> int:
        return len(self.buf)

    def put(self, b: bytes) -> int:
        self.buf += b"\0" * (-len(self.buf) % 8)
        at = len(self.buf)
        self.buf += b
        return at

    def dataset(self, arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr)
        data_at = self.put(arr.tobytes()) if arr.nbytes else UNDEF
        msgs = [_message(0x0001, _dataspace_message(arr.shape)),
                _message(0x0003, _dtype_message(arr.dtype), flags=1),            # constant message
                _message(0x0005, struct.pack("<BBBBI", 2, 2, 2, 1, 0)),          # fill value v2: late allocation, written if set, default (size 0)
                _message(0x0008, struct.pack("<BBQQ", 3, 1, data_at, arr.nbytes))]      # layout v3, contiguous
        return self.put(_object_header(msgs))

    def group(self, entries: Dict[str, Tuple[int, Optional[Tuple[int, int]]]], attrs: Dict[str, object]) -> Tuple[int, int, int]:
        """entries: name -> (object header address, (btree, heap) if the child is a group).  Returns (header, btree, heap)."""
        names = sorted(entries, key=lambda s: s.encode())
        heap_data = bytearray(8)                                   # offset 0: the empty name
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            heap_data += _pad8(n.encode() + b"\0")
        free = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 16)                     # one free block: next = 1 (none), size 16
        data_at = self.put(bytes(heap_data))
        heap_at = self.put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free, data_at))
        cap = 2 * self.leaf_k
        chunks = [names[i:i + cap] for i in range(0, len(names), cap)] or [[]]
        if len(chunks) > 2 * INTERNAL_K:
            raise ValueError("too many entries for a single-level group B-tree")
        snods = []
        for ch in chunks:
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(ch))
            for n in ch:
                addr, sub = entries[n]
                if sub is None:
                    body += struct.pack("<QQII16x", offs[n], addr, 0, 0)
                else:
                    body += struct.pack("<QQIIQQ", offs[n], addr, 1, 0, sub[0], sub[1])
            body += b"\0" * (8 + cap * 40 - len(body))
            snods.append(self.put(body))
        node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(chunks), UNDEF, UNDEF) + struct.pack("<Q", 0)
        for ch, at in zip(chunks, snods):
            node += struct.pack("<QQ", at, offs[ch[-1]] if ch else 0)
        node += b"\0" * (24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8 - len(node))
        btree_at = self.put(node)
        msgs = [_message(0x0011, struct.pack("<QQ", btree_at, heap_at))] + [_attribute_message(k, v) for k, v in attrs.items()]
        return self.put(_object_header(msgs)), btree_at, heap_at

    def finish(self, root: Tuple[int, int, int]) -> bytes:
        hdr, btree, heap = root
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.leaf_k, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQIIQQ", 0, hdr, 1, 0, btree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_bags(path, bags: Dict[str, Tuple[np.ndarray, np.ndarray, object]]) -> None:
    """One group per slide: 'feat' (stored as float16 like Step2_feature_extract.py:165), 'coords', attrs['label']."""
    leaf_k = max(4, -(-len(bags) // (2 * 2 * INTERNAL_K)))
    w = _Writer(leaf_k)
    root = {}
    for name, (feat, coords, label) in bags.items():
        f_at = w.dataset(np.asarray(feat).astype(np.float16))
        c_at = w.dataset(np.asarray(coords))
        g = w.group({"feat": (f_at, None), "coords": (c_at, None)}, {"label": np.asarray(label)})
        root[str(name)] = (g[0], (g[1], g[2]))
    data = w.finish(w.group(root, {}))
    with open(path, "wb") as fh:
        fh.write(data)


# ------------------------------------------------------------------------------------------ reader
class _Dataset:
    def __init__(self, mm, shape, dtype, offset, nbytes, inline=None):
        self._mm, self.shape, self.dtype, self._off, self._nbytes, self._inline = mm, shape, dtype, offset, nbytes, inline

    def __getitem__(self, key):
        if self._inline is not None:
            a = np.frombuffer(self._inline, dtype=self.dtype).reshape(self.shape)
        elif self._off == UNDEF or self._nbytes == 0:
            a = np.zeros(self.shape, self.dtype)
        else:
            a = np.frombuffer(self._mm, dtype=self.dtype, count=int(np.prod(self.shape, dtype=np.int64)), offset=self._off).reshape(self.shape)
        return np.array(a[key])          # a copy, like h5py's dataset[:]

    def view(self) -> np.ndarray:
        """Zero-copy view of a contiguous dataset (the memory-mapped file): what the prefetcher stages from."""
        return self[...] if self._inline is not None or self._off == UNDEF else np.frombuffer(
            self._mm, dtype=self.dtype, count=int(np.prod(self.shape, dtype=np.int64)), offset=self._off).reshape(self.shape)


class _Group:
    def __init__(self, f, header_at: int):
        self._f = f
        self.attrs, self._links, self._ds = f._read_object(header_at)

    def keys(self):
        return list(self._links)

    def __iter__(self):
        return iter(self._links)

    def __len__(self):
        return len(self._links)

    def __contains__(self, name):
        return name in self._links

    def __getitem__(self, name):
        obj = _Group(self._f, self._links[name])
        return obj._ds if obj._ds is not None else obj


class H5BagFile(_Group):
    """``h5py.File(path, 'r')`` for the subset described in the module docstring (own parser; needs no h5py)."""

    def __init__(self, path, mode: str = "r"):
        if mode != "r":
            raise ValueError("H5BagFile is read-only; use write_bags() to create a file")
        self._mm = np.memmap(path, dtype=np.uint8, mode="r")
        b = self._mm
        if bytes(b[:8]) != SIGNATURE:
            raise ValueError(f"{path}: not an HDF5 file (signature at offset 0 expected)")
        ver = int(b[8])
        if ver > 1:
            raise NotImplementedError("HDF5 superblock version >= 2 (libver='latest') is not supported: write with default settings")
        if int(b[13]) != 8 or int(b[14]) != 8:
            raise NotImplementedError("only 8-byte offsets / lengths are supported")
        o = 24 + (4 if ver == 1 else 0)
        base, _free, _eof, _drv = struct.unpack_from("<QQQQ", b, o)
        if base != 0:
            raise NotImplementedError("non-zero base address")
        _name_off, root_hdr = struct.unpack_from("<QQ", b, o + 32)
        super().__init__(self, root_hdr)

    def close(self):
        self._mm = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- object headers (version 1, with continuation blocks)
    def _messages(self, at: int):
        b = self._mm
        ver, _r, nmsg, _refs, size = struct.unpack_from("<BBHII", b, at)
        if ver != 1:
            raise NotImplementedError("version-2 object headers (libver='latest') are not supported")
        blocks = [(at + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", b, p)
                data = bytes(b[p + 8:p + 8 + msize])
                p += 8 + msize
                if flags & 2:
                    raise NotImplementedError("shared header messages are not supported")
                if mtype == 0x0010:                      # continuation: (address, length)
                    blocks.append(struct.unpack_from("<QQ", data, 0))
                out.append((mtype, data))
        return out

    def _read_object(self, at: int):
        attrs, links, shape, dtype, layout = {}, {}, None, None, None
        for mtype, d in self._messages(at):
            if mtype == 0x0011:                          # symbol table: old-style group
                btree, heap = struct.unpack_from("<QQ", d, 0)
                links = self._group_entries(btree, heap)
            elif mtype in (0x0002, 0x0006):
                raise NotImplementedError("new-style groups (link messages) are not supported: write with default settings")
            elif mtype == 0x0001:
                shape = _parse_dataspace(d)
            elif mtype == 0x0003:
                dtype = _parse_dtype(d)
            elif mtype == 0x0008:
                layout = d
            elif mtype == 0x000C:
                ver, _r, nlen, tlen, slen = struct.unpack_from("<BBHHH", d, 0)
                if ver != 1:
                    raise NotImplementedError(f"attribute message version {ver}")
                p = 8
                name = d[p:p + nlen].split(b"\0")[0].decode()
                p += nlen + (-nlen % 8)
                adt = _parse_dtype(d[p:p + tlen])
                p += tlen + (-tlen % 8)
                ashape = _parse_dataspace(d[p:p + slen])
                p += slen + (-slen % 8)
                cnt = int(np.prod(ashape, dtype=np.int64)) if ashape else 1
                val = np.frombuffer(d, dtype=adt, count=cnt, offset=p).reshape(ashape)
                attrs[name] = val[()] if ashape == () else val.copy()
        ds = None
        if layout is not None and shape is not None and dtype is not None:
            ver, cls = layout[0], layout[1]
            if ver != 3:
                raise NotImplementedError(f"data layout message version {ver}")
            if cls == 1:
                addr, nbytes = struct.unpack_from("<QQ", layout, 2)
                ds = _Dataset(self._mm, shape, dtype, addr, nbytes)
            elif cls == 0:
                (n,) = struct.unpack_from("<H", layout, 2)
                ds = _Dataset(self._mm, shape, dtype, 0, n, inline=layout[4:4 + n])
            else:
                raise NotImplementedError("chunked datasets are not supported (the reference writes contiguous ones)")
            ds.attrs = attrs
        return attrs, links, ds

    def _group_entries(self, btree: int, heap: int):
        b = self._mm
        if bytes(b[heap:heap + 4]) != b"HEAP":
            raise ValueError("corrupt file: local heap signature")
        _dsize, _free, data_at = struct.unpack_from("<QQQ", b, heap + 8)
        links = {}

        def name_at(off):
            p = data_at + off
            q = p
            while b[q] != 0:
                q += 1
            return bytes(b[p:q]).decode()

        def walk(node):
            sig = bytes(b[node:node + 4])
            if sig == b"TREE":
                _t, level, used = struct.unpack_from("<BBH", b, node + 4)
                for i in range(used):
                    (child,) = struct.unpack_from("<Q", b, node + 24 + 8 + 16 * i)
                    walk(child)
            elif sig == b"SNOD":
                (n,) = struct.unpack_from("<H", b, node + 6)
                for i in range(n):
                    noff, hdr = struct.unpack_from("<QQ", b, node + 8 + 40 * i)
                    links[name_at(noff)] = hdr
            else:
                raise ValueError("corrupt file: group B-tree node signature")

        walk(btree)
        return links


def open_bags(path):
    """h5py.File when h5py is available (any HDF5 flavour), else the built-in reader."""
    try:
        import h5py
        return h5py.File(path, "r")
    except ImportError:
        return H5BagFile(path)


# ------------------------------------------------------------------------------------------ the reference's loader
def split_dataset_camelyon(file_path, conf, split: Optional[dict] = None):
    """datasets/datasets.py:16-43: every slide's {'input': feat, 'coords', 'label'} by split.  ``split`` = the dict of the
    reference's ./splits/<dataset>/split_<seed>.json (train_names / val_names / test_names); without it, names containing
    'test' go to test and a seeded 10 % of the rest to validation (the reference uses sklearn's train_test_split there)."""
    h5 = open_bags(file_path)
    names = list(h5.keys())
    if split is None:
        test = [n for n in names if "test" in n]
        rest = [n for n in names if "test" not in n]
        rng = np.random.default_rng(int(getattr(conf, "seed", 0)))
        perm = rng.permutation(len(rest))
        n_val = max(1, int(round(0.1 * len(rest)))) if rest else 0
        val = [rest[i] for i in perm[:n_val]]
        train = [rest[i] for i in perm[n_val:]]
    else:
        train, val, test = split["train_names"], split["val_names"], split["test_names"]
    out = []
    for part in (train, val, test):
        d = {}
        for name in part:
            slide = h5[name]
            d[name] = {"input": slide["feat"][:], "coords": slide["coords"][:], "label": slide.attrs["label"]}
        out.append(d)
    h5.close()
    return out[0], train, out[1], val, out[2], test


class HDF5_feat_dataset2:
    """datasets/datasets.py:138-155."""

    def __init__(self, data_dict, data_names):
        self.data_dict = data_dict
        self.data_names = data_names

    def __len__(self):
        return len(self.data_names)

    def __getitem__(self, index):
        return self.data_dict[self.data_names[index]]


class BagPrefetcher:
    """Streams fp16 bags to the GPU ahead of the consumer: two pinned staging buffers and a copy stream, so the H2D copy of
    bag i + 1 overlaps the kernels of bag i.  Yields (name, x [N, D] float16 on the device, coords, label); the fp16 rows go
    to ``ACMIL_GA.forward_bags`` / ``forward`` as they are (the kernels widen them exactly)."""

    def __init__(self, dataset, device, max_rows: Optional[int] = None):
        import torch
        self.ds, self.dev = dataset, torch.device(device)
        items = [dataset[i] for i in range(len(dataset))]
        self.items = items
        rows = max_rows or max((it["input"].shape[0] for it in items), default=1)
        dim = items[0]["input"].shape[1] if items else 1
        self._pin = [torch.empty(rows, dim, dtype=torch.float16).pin_memory() for _ in range(2)]
        self._dev = [torch.empty(rows, dim, dtype=torch.float16, device=self.dev) for _ in range(2)]
        self._stream = torch.cuda.Stream(device=self.dev)
        self._ready = [torch.cuda.Event() for _ in range(2)]
        self._free = [torch.cuda.Event() for _ in range(2)]

    def _issue(self, i):
        import torch
        k = i % 2
        it = self.items[i]
        n = it["input"].shape[0]
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(self._free[k])          # the consumer of this buffer's previous bag has finished
            self._stream.synchronize()                       # ... so the pinned buffer may be overwritten by the host
            self._pin[k][:n].copy_(torch.from_numpy(np.ascontiguousarray(it["input"], dtype=np.float16)))
            self._dev[k][:n].copy_(self._pin[k][:n], non_blocking=True)
            self._ready[k].record(self._stream)

    def __iter__(self):
        import torch
        n_items = len(self.items)
        if n_items:
            self._issue(0)
        for i in range(n_items):
            if i + 1 < n_items:
                self._issue(i + 1)
            k = i % 2
            torch.cuda.current_stream(self.dev).wait_event(self._ready[k])
            it = self.items[i]
            name = self.ds.data_names[i] if hasattr(self.ds, "data_names") else i
            yield name, self._dev[k][:it["input"].shape[0]], it["coords"], it["label"]
            self._free[k].record(torch.cuda.current_stream(self.dev))
