"""Test infrastructure: composes small HDF5 files BYTE BY BYTE from the published file-format specification
("HDF5 File Format Specification Version 3.0"), in the two dialects a NetCDF-4 file can come in:

  style "old"  superblock 0, version-1 object headers, the root group as a symbol table (v1 B-tree of SNOD nodes over a
               local heap) -- what libhdf5 writes with libver "earliest" and no creation-order tracking (h5py defaults,
               MATLAB v7.3);
  style "new"  superblock 2, version-2 object headers ("OHDR", optional time stamps / creation order / continuation
               "OCHK" blocks), links as link messages in the root header, or -- beyond eight objects -- "dense" in a
               fractal heap (root direct block, or a root indirect block over several direct blocks): what libnetcdf's
               creation-order tracking produces.

Datasets: float / double of either byte order; compact, contiguous or chunked storage (v1 chunk B-tree of one or two
levels, edge chunks stored whole as HDF5 does) with the filter pipeline deflate / shuffle / fletcher32.

Neither libhdf5 nor libnetcdf exists in this image, so nothing here is checked against them; the one real HDF5 file the
image holds (scipy's MATLAB v7.3 sample) pins the "old" dialect of the reader independently of this writer.  Nothing
under ampe_b200/ imports this module.
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _rot(x, k):
    return ((x << k) | (x >> (32 - k))) & 0xFFFFFFFF


def lookup3(data, init=0):
    """Bob Jenkins' hashlittle: the metadata checksum of the version-2 structures"""
    n = len(data)
    a = b = c = (0xDEADBEEF + n + init) & 0xFFFFFFFF
    p = 0
    M = 0xFFFFFFFF
    while n > 12:
        a = (a + int.from_bytes(data[p:p + 4], "little")) & M
        b = (b + int.from_bytes(data[p + 4:p + 8], "little")) & M
        c = (c + int.from_bytes(data[p + 8:p + 12], "little")) & M
        a = (a - c) & M; a ^= _rot(c, 4); c = (c + b) & M
        b = (b - a) & M; b ^= _rot(a, 6); a = (a + c) & M
        c = (c - b) & M; c ^= _rot(b, 8); b = (b + a) & M
        a = (a - c) & M; a ^= _rot(c, 16); c = (c + b) & M
        b = (b - a) & M; b ^= _rot(a, 19); a = (a + c) & M
        c = (c - b) & M; c ^= _rot(b, 4); b = (b + a) & M
        p += 12
        n -= 12
    if n == 0:
        return c
    tail = data[p:] + b"\0" * (12 - n)
    a = (a + int.from_bytes(tail[0:4], "little")) & M
    b = (b + int.from_bytes(tail[4:8], "little")) & M
    c = (c + int.from_bytes(tail[8:12], "little")) & M
    c ^= b; c = (c - _rot(b, 14)) & M
    a ^= c; a = (a - _rot(c, 11)) & M
    b ^= a; b = (b - _rot(a, 25)) & M
    c ^= b; c = (c - _rot(b, 16)) & M
    a ^= c; a = (a - _rot(c, 4)) & M
    b ^= a; b = (b - _rot(a, 14)) & M
    c ^= b; c = (c - _rot(b, 24)) & M
    return c


def fletcher32(data):
    s1 = s2 = 0
    n = len(data) // 2
    for i in range(n):
        s1 = (s1 + ((data[2 * i] << 8) | data[2 * i + 1])) % 65535
        s2 = (s2 + s1) % 65535
    if len(data) & 1:
        s1 = (s1 + (data[-1] << 8)) % 65535
        s2 = (s2 + s1) % 65535
    return (s2 << 16) | s1


class _Image:
    """the file under construction: append-only allocation with back-patching"""

    def __init__(self, reserve):
        self.b = bytearray(reserve)

    def alloc(self, data, align=8):
        while len(self.b) % align:
            self.b.append(0)
        a = len(self.b)
        self.b += data
        return a

    def reserve(self, n, align=8):
        return self.alloc(bytes(n), align)

    def put(self, addr, data):
        self.b[addr:addr + len(data)] = data


def _pad8(b):
    return b + bytes(-len(b) % 8)


def _datatype(dtype):
    dt = np.dtype(dtype)
    be = 1 if dt.byteorder == ">" else 0
    if dt.kind != "f" or dt.itemsize not in (4, 8):
        raise ValueError("float32 / float64 only")
    if dt.itemsize == 8:
        return bytes([0x11, 0x20 | be, 63, 0]) + struct.pack("<IHHBBBBI", 8, 0, 64, 52, 11, 0, 52, 1023)
    return bytes([0x11, 0x20 | be, 31, 0]) + struct.pack("<IHHBBBBI", 4, 0, 32, 23, 8, 0, 23, 127)


def _dataspace(shape, version, with_max=False):
    flags = 1 if with_max else 0
    if version == 1:
        b = bytes([1, len(shape), flags, 0, 0, 0, 0, 0])
    else:
        b = bytes([2, len(shape), flags, 1])
    b += b"".join(struct.pack("<Q", int(s)) for s in shape)
    if with_max:
        b += b"".join(struct.pack("<Q", int(s)) for s in shape)
    return b


def _filters_msg(filters, version, itemsize):
    out = bytes([1, len(filters), 0, 0, 0, 0, 0, 0]) if version == 1 else bytes([2, len(filters)])
    for name in filters:
        fid, cd = {"deflate": (1, [4]), "shuffle": (2, [itemsize]), "fletcher32": (3, [])}[name]
        if version == 1:
            nm = _pad8(name.encode() + b"\0")
            out += struct.pack("<HHHH", fid, len(nm), 1, len(cd)) + nm
            out += b"".join(struct.pack("<I", v) for v in cd)
            if len(cd) & 1:
                out += bytes(4)
        else:
            out += struct.pack("<HHH", fid, 1, len(cd)) + b"".join(struct.pack("<I", v) for v in cd)
    return out


def _attribute_v1(name, value):
    """a scalar int32 attribute (libnetcdf's _Netcdf4Dimid): the reader must step over it"""
    nm = name.encode() + b"\0"
    dt = bytes([0x10, 0x08, 0, 0]) + struct.pack("<IHH", 4, 0, 32)
    ds = bytes([1, 0, 0, 0, 0, 0, 0, 0])
    return bytes([1, 0]) + struct.pack("<HHH", len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + struct.pack("<i", value)


def _apply_filters(raw, filters, itemsize):
    for name in filters:
        if name == "shuffle":
            a = np.frombuffer(raw, dtype=np.uint8).reshape(-1, itemsize)
            raw = a.T.copy().tobytes()
        elif name == "deflate":
            raw = zlib.compress(raw, 4)
        elif name == "fletcher32":
            raw = raw + struct.pack("<I", fletcher32(raw))
    return raw


def _store_dataset(img, arr, layout, chunks, filters, leaf_fanout, skip_deflate_on_first):
    """writes the raw data (and the chunk index); returns the layout message (version 3)"""
    arr = np.ascontiguousarray(arr)
    if layout == "compact":
        raw = arr.tobytes()
        return bytes([3, 0]) + struct.pack("<H", len(raw)) + raw
    if layout == "contiguous":
        raw = arr.tobytes()
        return bytes([3, 1]) + struct.pack("<QQ", img.alloc(raw), len(raw))
    rank = arr.ndim
    recs = []
    import itertools
    for origin in itertools.product(*[range(0, arr.shape[d], chunks[d]) for d in range(rank)]):
        block = np.zeros(chunks, dtype=arr.dtype)  # edge chunks are stored whole
        sl = tuple(slice(origin[d], min(origin[d] + chunks[d], arr.shape[d])) for d in range(rank))
        block[tuple(slice(0, s.stop - s.start) for s in sl)] = arr[sl]
        mask = 0
        use = list(filters)
        if skip_deflate_on_first and not recs and "deflate" in use:
            mask = 1 << use.index("deflate")  # a chunk the writer left uncompressed: filter mask bit set
            use = [f for f in use if f != "deflate"]
        raw = _apply_filters(block.tobytes(), use, arr.dtype.itemsize)
        recs.append((origin, len(raw), mask, img.alloc(raw)))

    def key(nbytes, mask, origin):
        return struct.pack("<II", nbytes, mask) + b"".join(struct.pack("<Q", o) for o in origin) + struct.pack("<Q", 0)

    end_key = key(0, 0, [(-(-arr.shape[d] // chunks[d])) * chunks[d] for d in range(rank)])

    def node(level, entries):
        body = b"TREE" + bytes([1, level]) + struct.pack("<HQQ", len(entries), UNDEF, UNDEF)
        for k, child in entries:
            body += k + struct.pack("<Q", child)
        return img.alloc(body + end_key)

    leaves = []
    for i in range(0, len(recs), leaf_fanout):
        part = recs[i:i + leaf_fanout]
        leaves.append((key(part[0][1], part[0][2], part[0][0]), node(0, [(key(nb, m, o), a) for o, nb, m, a in part])))
    root = leaves[0][1] if len(leaves) == 1 else node(1, leaves)
    msg = bytes([3, 2, rank + 1]) + struct.pack("<Q", root)
    msg += b"".join(struct.pack("<I", c) for c in chunks) + struct.pack("<I", arr.dtype.itemsize)
    return msg


def _header_v1(img, msgs, continuation_after=None):
    def enc(ms):
        return b"".join(struct.pack("<HHB3x", t, len(_pad8(d)), 0) + _pad8(d) for t, d in ms)

    if continuation_after is None or continuation_after >= len(msgs):
        body = enc(msgs)
        return img.alloc(struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body)
    tail = enc(msgs[continuation_after:])
    tail_addr = img.alloc(tail)
    first = msgs[:continuation_after] + [(0x10, struct.pack("<QQ", tail_addr, len(tail)))]
    body = enc(first)
    return img.alloc(struct.pack("<BBHII4x", 1, 0, len(msgs) + 1, 1, len(body)) + body)


def _header_v2(img, msgs, times, track_order, continuation_after=None):
    flags = (0x20 if times else 0) | (0x04 | 0x10 if track_order else 0) | 1  # chunk-0 size in two bytes

    def enc(ms):
        out = b""
        for i, (t, d) in enumerate(ms):
            out += struct.pack("<BHB", t, len(d), 0) + (struct.pack("<H", i) if track_order else b"") + d
        return out

    def prefix(n):
        p = b"OHDR" + bytes([2, flags])
        if times:
            p += struct.pack("<IIII", 1700000000, 1700000000, 1700000000, 1700000000)
        if track_order:
            p += struct.pack("<HH", 8, 6)
        return p + struct.pack("<H", n)

    if continuation_after is None or continuation_after >= len(msgs):
        body = enc(msgs) + bytes(3)  # a gap too small for a message header
        blk = prefix(len(body)) + body
        return img.alloc(blk + struct.pack("<I", lookup3(blk)))
    tail = b"OCHK" + enc(msgs[continuation_after:])
    tail += struct.pack("<I", lookup3(tail))
    tail_addr = img.alloc(tail)
    body = enc(msgs[:continuation_after] + [(0x10, struct.pack("<QQ", tail_addr, len(tail)))])
    blk = prefix(len(body)) + body
    return img.alloc(blk + struct.pack("<I", lookup3(blk)))


def _link_msg(name, addr, order):
    nm = name.encode()
    return bytes([1, 0x04 | 0x10]) + struct.pack("<Q", order) + bytes([0]) + bytes([len(nm)]) + nm + struct.pack("<Q", addr)


def _dense_links(img, links, indirect_root):
    """link messages as managed objects of a fractal heap; returns the link-info message"""
    width, start, max_direct, heap_bits = 4, 512, 65536, 32
    hdr_size = 22 + 12 * 8 + 3 * 8 + 4
    hdr = img.reserve(hdr_size)
    objs = [_link_msg(n, a, i) for i, (n, a) in enumerate(links)]
    head = 5 + 8 + heap_bits // 8 + 4  # signature, version, heap header address, block offset, checksum

    def direct(block_off, size, payload):
        blk = b"FHDB" + bytes([0]) + struct.pack("<Q", hdr) + struct.pack("<I", block_off)
        blk += bytes(4) + payload
        assert len(blk) <= size, "test writer: too many links for one direct block"
        blk += bytes(size - len(blk))
        blk = blk[:head - 4] + struct.pack("<I", lookup3(blk[:head - 4] + bytes(4) + blk[head:])) + blk[head:]
        return img.alloc(blk)

    if not indirect_root:
        root, rows = direct(0, start, b"".join(objs)), 0
    else:
        # rows 0 and 1 hold blocks of the starting size, row 2 twice that: three blocks, the entries between them unused
        third = (len(objs) + 2) // 3
        parts = [objs[:third], objs[third:2 * third], objs[2 * third:]]
        slots = [0, 1, 2 * width + 1]  # row 0 col 0, row 0 col 1, row 2 col 1
        entries = [UNDEF] * (3 * width)
        for part, slot in zip(parts, slots):
            r, col = divmod(slot, width)
            size = start if r < 2 else start << (r - 1)
            off = sum((start if rr < 2 else start << (rr - 1)) * width for rr in range(r)) + col * size
            entries[slot] = direct(off, size, b"".join(part))
        blk = b"FHIB" + bytes([0]) + struct.pack("<Q", hdr) + struct.pack("<I", 0) + b"".join(struct.pack("<Q", e) for e in entries)
        root, rows = img.alloc(blk + struct.pack("<I", lookup3(blk))), 3
    h = b"FRHP" + bytes([0]) + struct.pack("<HHB", 7, 0, 2) + struct.pack("<I", 4096)
    h += struct.pack("<QQQQ", 0, UNDEF, 0, UNDEF)
    h += struct.pack("<QQQQ", start, start, start, len(objs))
    h += struct.pack("<QQQQ", 0, 0, 0, 0)
    h += struct.pack("<HQQHH", width, start, max_direct, heap_bits, 1) + struct.pack("<QH", root, rows)
    h += struct.pack("<I", lookup3(h))
    assert len(h) == hdr_size
    img.put(hdr, h)
    # the name / creation-order indices (version-2 B-trees) are not written: the reader walks the heap blocks
    return bytes([0, 3]) + struct.pack("<QQQQ", len(objs), hdr, UNDEF, UNDEF)


def write_hdf5(path, variables, style="new", dense=None, indirect_root=False, times=True, track_order=True,
               continuation=False, userblock=0, dimensions=None):
    """variables: {name: array} or {name: (array, options)}; options: layout "contiguous" | "compact" | "chunked",
    chunks, filters [..."shuffle", "deflate", "fletcher32"], leaf_fanout, skip_deflate_on_first, dataspace_version,
    filter_version.  dimensions: {name: size} written the way libnetcdf writes a dimension without a variable (a
    one-dimensional big-endian float dataset with no storage allocated)."""
    img = _Image(96 if style == "old" else 48)
    items = []
    for name, spec in variables.items():
        arr, opt = spec if isinstance(spec, tuple) else (spec, {})
        arr = np.asarray(arr)
        layout = opt.get("layout", "contiguous")
        filters = opt.get("filters", [])
        lay = _store_dataset(img, arr, layout, opt.get("chunks"), filters, opt.get("leaf_fanout", 64),
                             opt.get("skip_deflate_on_first", False))
        msgs = [(0x01, _dataspace(arr.shape, opt.get("dataspace_version", 1 if style == "old" else 2), opt.get("with_max", False))),
                (0x03, _datatype(arr.dtype)), (0x05, bytes([2, 2, 0, 0]))]
        if filters:
            msgs.append((0x0B, _filters_msg(filters, opt.get("filter_version", 1 if style == "old" else 2), arr.dtype.itemsize)))
        msgs += [(0x08, lay), (0x0C, _attribute_v1("_Netcdf4Dimid", 0))]
        items.append((name, msgs))
    for name, size in (dimensions or {}).items():
        msgs = [(0x01, _dataspace((size,), 1 if style == "old" else 2, True)), (0x03, _datatype(">f4")),
                (0x05, bytes([2, 2, 0, 0])), (0x08, bytes([3, 1]) + struct.pack("<QQ", UNDEF, 4 * size))]
        items.append((name, msgs))
    cont = 2 if continuation else None
    links = []
    for name, msgs in items:
        if style == "old":
            links.append((name, _header_v1(img, msgs, cont)))
        else:
            links.append((name, _header_v2(img, msgs, times, track_order, cont)))
    if style == "old":
        names = sorted(links)
        seg = bytearray(8)
        offs = {}
        for n, _ in names:
            offs[n] = len(seg)
            seg += _pad8(n.encode() + b"\0")
        seg_addr = img.alloc(bytes(seg))
        heap = img.alloc(b"HEAP" + bytes(4) + struct.pack("<QQQ", len(seg), UNDEF, seg_addr))
        entries = []
        for i in range(0, len(names), 3):  # several symbol-table nodes
            part = names[i:i + 3]
            snod = b"SNOD" + bytes([1, 0]) + struct.pack("<H", len(part))
            for n, a in part:
                snod += struct.pack("<QQII16x", offs[n], a, 0, 0)
            entries.append((offs[part[-1][0]], img.alloc(snod)))
        tree = b"TREE" + bytes([0, 0]) + struct.pack("<HQQ", len(entries), UNDEF, UNDEF) + struct.pack("<Q", 0)
        for k, child in entries:
            tree += struct.pack("<QQ", child, k)
        tree_addr = img.alloc(tree)
        root = _header_v1(img, [(0x11, struct.pack("<QQ", tree_addr, heap))])
        sb = b"\x89HDF\r\n\x1a\n" + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", 4, 16, 0)
        sb += struct.pack("<QQQQ", userblock, UNDEF, len(img.b), UNDEF)
        sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", tree_addr, heap)
        assert len(sb) == 96
    else:
        use_dense = dense if dense is not None else len(links) > 8
        gmsgs = []
        if use_dense:
            gmsgs.append((0x02, _dense_links(img, links, indirect_root)))
        else:
            gmsgs.append((0x02, bytes([0, 3]) + struct.pack("<QQQQ", len(links), UNDEF, UNDEF, UNDEF)))
            gmsgs += [(0x06, _link_msg(n, a, i)) for i, (n, a) in enumerate(links)]
        gmsgs.append((0x0A, bytes([0, 0])))
        root = _header_v2(img, gmsgs, times, track_order, 3 if continuation and not use_dense else None)
        sb = b"\x89HDF\r\n\x1a\n" + bytes([2, 8, 8, 0]) + struct.pack("<QQQQ", userblock, UNDEF, len(img.b), root)
        sb += struct.pack("<I", lookup3(sb))
        assert len(sb) == 48
    img.put(0, sb)
    with open(path, "wb") as f:
        f.write(bytes(userblock) + bytes(img.b))
