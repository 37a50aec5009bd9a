"""Writes the netCDF-4 (HDF5) fixtures tests/golden/dataset_nc4_*.nc -- a DSSTNE dataset file in the schema of
U/NetCDFhelper.cpp:332-372 (analog sparse data set) laid out the way the HDF5 library lays out a netCDF-4 file, byte by byte
from the "HDF5 File Format Specification Version 3.0" (no libhdf5 / netCDF4 / h5py exists in this image, and the reference's own
.nc blobs are missing from its tree):

  dataset_nc4_old.nc   superblock v0, version-1 object headers (with a continuation block), root group as a symbol table
                       (B-tree v1 + SNOD + local heap), dataspace v1, attribute messages v1 in the headers, contiguous layout v3 --
                       what libhdf5 writes with the default (earliest) format bounds
  dataset_nc4_new.nc   superblock v2, version-2 object headers ("OHDR", times, creation order, an "OCHK" continuation), link and
                       attribute DENSE storage in fractal heaps (a root direct block for the links, a root INDIRECT block with two
                       direct blocks for the global attributes), dataspace v2, attribute messages v3, layouts v3 contiguous,
                       v3 compact and v4 contiguous, one big-endian variable -- what netcdf-c >= 4.x writes with creation-order
                       tracking on and more than eight links / attributes per object
Both carry the netCDF-4 bookkeeping the reader must drop: dimension-scale data sets with CLASS / NAME / _Netcdf4Dimid, _NCProperties,
a DIMENSION_LIST attribute of variable-length type.  Metadata checksums are Jenkins lookup3, as the library computes them.

    python tests/golden/make_hdf5_fixture.py        # rewrites the two files next to this script
"""
import os
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
HERE = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------- Jenkins lookup3 hashlittle (the HDF5 metadata checksum)
def _rot(x, k):
    return ((x << k) | (x >> (32 - k))) & 0xFFFFFFFF


def lookup3(data, init=0):
    a = b = c = (0xDEADBEEF + len(data) + init) & 0xFFFFFFFF
    p, n = 0, len(data)
    while n > 12:
        a = (a + int.from_bytes(data[p:p + 4], "little")) & 0xFFFFFFFF
        b = (b + int.from_bytes(data[p + 4:p + 8], "little")) & 0xFFFFFFFF
        c = (c + int.from_bytes(data[p + 8:p + 12], "little")) & 0xFFFFFFFF
        a = (a - c) & 0xFFFFFFFF; a ^= _rot(c, 4); c = (c + b) & 0xFFFFFFFF
        b = (b - a) & 0xFFFFFFFF; b ^= _rot(a, 6); a = (a + c) & 0xFFFFFFFF
        c = (c - b) & 0xFFFFFFFF; c ^= _rot(b, 8); b = (b + a) & 0xFFFFFFFF
        a = (a - c) & 0xFFFFFFFF; a ^= _rot(c, 16); c = (c + b) & 0xFFFFFFFF
        b = (b - a) & 0xFFFFFFFF; b ^= _rot(a, 19); a = (a + c) & 0xFFFFFFFF
        c = (c - b) & 0xFFFFFFFF; c ^= _rot(b, 4); b = (b + a) & 0xFFFFFFFF
        p += 12; n -= 12
    if n == 0:
        return c
    tail = bytes(data[p:p + n]) + b"\0" * (12 - n)
    a = (a + int.from_bytes(tail[0:4], "little")) & 0xFFFFFFFF
    b = (b + int.from_bytes(tail[4:8], "little")) & 0xFFFFFFFF
    c = (c + int.from_bytes(tail[8:12], "little")) & 0xFFFFFFFF
    c ^= b; c = (c - _rot(b, 14)) & 0xFFFFFFFF
    a ^= c; a = (a - _rot(c, 11)) & 0xFFFFFFFF
    b ^= a; b = (b - _rot(a, 25)) & 0xFFFFFFFF
    c ^= b; c = (c - _rot(b, 16)) & 0xFFFFFFFF
    a ^= c; a = (a - _rot(c, 4)) & 0xFFFFFFFF
    b ^= a; b = (b - _rot(a, 14)) & 0xFFFFFFFF
    c ^= b; c = (c - _rot(b, 24)) & 0xFFFFFFFF
    return c


# ---------------------------------------------------------------- file image
class Image:
    def __init__(self):
        self.buf = bytearray()

    def alloc(self, n, align=8):
        while len(self.buf) % align:
            self.buf.append(0)
        off = len(self.buf)
        self.buf.extend(b"\0" * n)
        return off

    def put(self, off, data):
        self.buf[off:off + len(data)] = data

    def add(self, data, align=8):
        off = self.alloc(len(data), align)
        self.put(off, data)
        return off


def pad8(b):
    return b + b"\0" * (-len(b) % 8)


# ---------------------------------------------------------------- messages
def datatype(dt, big=False):
    """dt: numpy dtype, or ("str", n) for a fixed-length string"""
    if isinstance(dt, tuple):
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, dt[1])
    dt = np.dtype(dt)
    order = 1 if big else 0
    if dt.kind in "ui":
        return struct.pack("<BBBBIHH", 0x10, order | (0x08 if dt.kind == "i" else 0), 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt == np.float32:
        return struct.pack("<BBBBIHHBBBBI", 0x11, order | 0x20, 31, 0, 4, 0, 32, 23, 8, 0, 23, 127)
    if dt == np.float64:
        return struct.pack("<BBBBIHHBBBBI", 0x11, order | 0x20, 63, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
    raise ValueError(dt)


def vlen_reference_type():
    """a variable-length sequence of object references (the type of DIMENSION_LIST): class 9 over class 7"""
    base = struct.pack("<BBBBI", 0x17, 0x00, 0, 0, 8)
    return struct.pack("<BBBBI", 0x19, 0x00, 0, 0, 16) + base


def dataspace(dims, version):
    if version == 1:
        return struct.pack("<BBBB4x", 1, len(dims), 0, 0) + b"".join(struct.pack("<Q", d) for d in dims)
    return struct.pack("<BBBB", 2, len(dims), 0, 1 if dims else 0) + b"".join(struct.pack("<Q", d) for d in dims)


def attribute(name, dtmsg, dims, data, version):
    nm = name.encode() + b"\0"
    ds = dataspace(dims, 1 if version == 1 else 2)
    if version == 1:
        return struct.pack("<BBHHH", 1, 0, len(nm), len(dtmsg), len(ds)) + pad8(nm) + pad8(dtmsg) + pad8(ds) + data
    return struct.pack("<BBHHHB", 3, 0, len(nm), len(dtmsg), len(ds), 0) + nm + dtmsg + ds + data


def att_uint(name, value, version):
    return attribute(name, datatype(np.uint32), [], struct.pack("<I", value), version)


def att_text(name, text, version):
    b = text.encode()
    return attribute(name, datatype(("str", len(b))), [], b, version)


def layout_contiguous(addr, size, version=3):
    return struct.pack("<BBQQ", version, 1, addr, size)


def layout_compact(data):
    return struct.pack("<BBH", 3, 0, len(data)) + data


# ---------------------------------------------------------------- object headers
def header_v1(img, msgs, split_after=None):
    """msgs: [(type, bytes)].  split_after: put the messages after this index into a continuation block."""
    def enc(ms):
        return b"".join(struct.pack("<HHB3x", t, len(pad8(d)), 0) + pad8(d) for t, d in ms)
    first, rest = (msgs, []) if split_after is None else (msgs[:split_after], msgs[split_after:])
    n = len(msgs) + (1 if rest else 0)
    body = enc(first)
    cont_msg_off = None
    if rest:
        cont_msg_off = len(body) + 8
        body += struct.pack("<HHB3x", 0x10, 16, 0) + b"\0" * 16
    off = img.alloc(16 + len(body))
    img.put(off, struct.pack("<BBHII4x", 1, 0, n, 1, len(body)) + body)
    if rest:
        blk = enc(rest)
        coff = img.add(blk)
        img.put(off + 16 + cont_msg_off, struct.pack("<QQ", coff, len(blk)))
    return off


def header_v2(img, msgs, split_after=None):
    flags = 0x20 | 0x04 | 0x02                                    # times stored, attribute creation order tracked, 4-byte chunk size
    def enc(ms, start):
        return b"".join(struct.pack("<BHBH", t, len(d), 0, start + i) + d for i, (t, d) in enumerate(ms))
    first, rest = (msgs, []) if split_after is None else (msgs[:split_after], msgs[split_after:])
    body = enc(first, 0)
    cont_at = None
    if rest:
        cont_at = len(body) + 6
        body += struct.pack("<BHBH", 0x10, 16, 0, 0) + b"\0" * 16
    prefix = b"OHDR" + struct.pack("<BB", 2, flags) + struct.pack("<IIII", 1500000000, 1500000000, 1500000000, 1500000000) + struct.pack("<I", len(body))
    off = img.alloc(len(prefix) + len(body) + 4)
    if rest:
        blk = b"OCHK" + enc(rest, len(first))
        blk += struct.pack("<I", lookup3(blk))
        coff = img.add(blk)
        body = body[:cont_at] + struct.pack("<QQ", coff, len(blk)) + body[cont_at + 16:]
    whole = prefix + body
    img.put(off, whole + struct.pack("<I", lookup3(whole)))
    return off


# ---------------------------------------------------------------- fractal heap holding `objects` (bytes each), link or attribute messages
def fractal_heap(img, objects, start_size, indirect):
    """Managed objects packed block by block in heap order.  indirect = False: the root IS a direct block (everything must fit it);
    True: a root indirect block over a doubling table of width 4 (rows 0 and 1: start_size, row r >= 2: start_size * 2^(r-1))."""
    width, max_direct, heap_bits = 4, 65536, 32
    hdr_len = 4 + 1 + 2 + 2 + 1 + 4 + 8 + 8 + 8 + 8 + 8 * 8 + 2 + 8 + 8 + 2 + 2 + 8 + 2 + 4
    hoff = img.alloc(hdr_len)
    dhdr = 5 + 8 + 4 + 4

    def block_size(i):
        row = i // width
        return start_size << (row - 1 if row > 1 else 0)
    blocks, cur = [], bytearray()
    for o in objects:                                             # an object never straddles two blocks
        while dhdr + len(cur) + len(o) > block_size(len(blocks)):
            assert cur or dhdr + len(o) <= max_direct, "object larger than any direct block"
            blocks.append(cur); cur = bytearray()
        cur += o
    blocks.append(cur)
    assert indirect or len(blocks) == 1, "objects do not fit the root direct block"
    addrs, heap_off = [], 0
    for i, payload in enumerate(blocks):
        size = block_size(i)
        assert size <= max_direct
        boff = img.alloc(size)
        head = b"FHDB" + struct.pack("<BQI", 0, hoff, heap_off)
        blk = bytearray(head + b"\0\0\0\0" + payload)
        blk += b"\0" * (size - len(blk))
        blk[len(head):len(head) + 4] = struct.pack("<I", lookup3(bytes(blk[:len(head)]) + b"\0\0\0\0" + bytes(blk[len(head) + 4:])))
        img.put(boff, bytes(blk))
        addrs.append(boff)
        heap_off += size
    if indirect:
        rows = (len(blocks) + width - 1) // width
        ib = b"FHIB" + struct.pack("<BQI", 0, hoff, 0) + b"".join(struct.pack("<Q", addrs[i] if i < len(addrs) else UNDEF) for i in range(width * rows))
        root = img.add(ib + struct.pack("<I", lookup3(ib)))
    else:
        root, rows = addrs[0], 0
    nobj = len(objects)
    h = (b"FRHP" + struct.pack("<BHHBI", 0, 7, 0, 0x02, 4096) + struct.pack("<QQQQ", 0, UNDEF, 0, UNDEF) +
         struct.pack("<QQQQQQQQ", heap_off, heap_off, heap_off, nobj, 0, 0, 0, 0) +
         struct.pack("<HQQHHQH", width, start_size, max_direct, heap_bits, 1, root, rows))
    img.put(hoff, h + struct.pack("<I", lookup3(h)))
    return hoff


def link_message(name, addr, order):
    nm = name.encode()
    return struct.pack("<BBQB", 1, 0x04, order, len(nm)) + nm + struct.pack("<Q", addr)      # creation order present, 1-byte name length, hard link


# ---------------------------------------------------------------- the data set (U/NetCDFhelper.cpp:332-372)
def sample():
    start = np.array([0, 3, 3, 7, 9], dtype=np.uint32)
    end = np.array([3, 3, 7, 9, 12], dtype=np.uint32)
    index = np.array([5, 17, 130, 2, 3, 64, 255, 9, 100, 0, 31, 200], dtype=np.uint32)
    data = (np.arange(12, dtype=np.float32) * 0.25 + 0.5).astype(np.float32)
    gatts = [("datasets", 1), ("name0", "gl_input"), ("attributes0", 1), ("kind0", 0), ("dataType0", 4), ("dimensions0", 1), ("width0", 256)]
    return start, end, index, data, gatts


def dimension_scale(img, name, size, dimid, hdr, av):
    msgs = [(0x01, dataspace([size], 1 if av == 1 else 2)), (0x03, datatype(np.float32, big=True)), (0x08, layout_contiguous(UNDEF, 0)),
            (0x0C, att_text("CLASS", "DIMENSION_SCALE\0", av)),
            (0x0C, att_text("NAME", "This is a netCDF dimension but not a netCDF variable.%10d\0" % size, av)),
            (0x0C, attribute("_Netcdf4Dimid", datatype(np.int32), [], struct.pack("<i", dimid), av))]
    return hdr(img, msgs)


def fletcher32(b):
    """HDF5's checksum of a chunk (H5_checksum_fletcher32): 16-bit big-endian words, both sums folded modulo 65535"""
    s1 = s2 = 0
    for i in range(0, len(b) - 1, 2):
        s1 = (s1 + ((b[i] << 8) | b[i + 1])) % 65535
        s2 = (s2 + s1) % 65535
    if len(b) % 2:
        s1 = (s1 + (b[-1] << 8)) % 65535
        s2 = (s2 + s1) % 65535
    return (s2 << 16) | s1


def filter_pipeline(filters, elem, version):
    """message 0x0B: filters = names out of shuffle / deflate / fletcher32 in pipeline order"""
    ids = {"deflate": (1, [6]), "shuffle": (2, [elem]), "fletcher32": (3, [])}
    out = struct.pack("<BB6x", 1, len(filters)) if version == 1 else struct.pack("<BB", 2, len(filters))
    for f in filters:
        fid, vals = ids[f]
        if version == 1:
            nm = pad8(f.encode() + b"\0")
            out += struct.pack("<HHHH", fid, len(nm), 1 if f != "fletcher32" else 0, len(vals)) + nm + b"".join(struct.pack("<I", x) for x in vals)
            if len(vals) % 2:
                out += b"\0" * 4
        else:
            out += struct.pack("<HHH", fid, 1 if f != "fletcher32" else 0, len(vals)) + b"".join(struct.pack("<I", x) for x in vals)
    return out


def chunk_btree(img, entries, end, fanout=64):
    """version-1 B-tree of node type 1 over 1-D chunks; entries = [(start element, stored bytes, filter mask, address)], in order"""
    def key(size, mask, start):
        start = start if isinstance(start, tuple) else (start,)
        return struct.pack("<II", size, mask) + b"".join(struct.pack("<Q", x) for x in start) + struct.pack("<Q", 0)

    def node(level, items, upper):                                 # items = [(first start, size, mask, child address)]
        body = b"TREE" + struct.pack("<BBHQQ", 1, level, len(items), UNDEF, UNDEF)
        for start, size, mask, child in items:
            body += key(size, mask, start) + struct.pack("<Q", child)
        body += key(0, 0, upper)
        body += b"\0" * max(0, 24 + (2 * 32 + 1) * len(key(0, 0, upper)) + 2 * 32 * 8 - len(body))    # nodes are allocated for 2 K = 64 children
        return img.add(body)
    level, items = 0, entries
    while True:
        groups = [items[i:i + fanout] for i in range(0, len(items), fanout)] or [[]]
        nodes = []
        for gi, g in enumerate(groups):
            upper = groups[gi + 1][0][0] if gi + 1 < len(groups) else end
            nodes.append((g[0][0] if g else 0, g[0][1] if g else 0, g[0][2] if g else 0, node(level, g, upper)))
        if len(nodes) == 1:
            return nodes[0][3]
        level, items = level + 1, nodes


def variable(img, name, arr, hdr, av, how="contiguous", big=False, chunk=0, filters=(), skip_filter_on=None, holes=()):
    """how = contiguous | compact | v4 (layout message version 4) | chunked (chunk elements per chunk, `filters` applied in order;
    chunk number skip_filter_on is stored with its deflate step skipped and flagged in its mask; chunk numbers in `holes` are never written)"""
    raw = arr.astype(arr.dtype.newbyteorder(">" if big else "<")).tobytes()
    extra = []
    if how == "compact":
        lay = layout_compact(raw)
    elif how == "chunked":
        import zlib
        elem = arr.dtype.itemsize
        entries = []
        if arr.ndim == 2:                                           # chunk = (rows, columns); chunks in row-major order of their origin
            origins = [(r0, c0) for r0 in range(0, arr.shape[0], chunk[0]) for c0 in range(0, arr.shape[1], chunk[1])]
            stored = arr.astype(arr.dtype.newbyteorder(">" if big else "<"))
        else:
            origins = list(range(0, len(arr), chunk))
        for ci, first in enumerate(origins):
            if ci in holes:
                continue
            if arr.ndim == 2:
                tile = np.zeros(chunk, dtype=stored.dtype)
                part = stored[first[0]:first[0] + chunk[0], first[1]:first[1] + chunk[1]]
                tile[:part.shape[0], :part.shape[1]] = part
                piece = tile.tobytes()
            else:
                piece = raw[first * elem:(first + chunk) * elem]
                piece += b"\0" * (chunk * elem - len(piece))        # edge chunks are stored whole
            mask = 0
            for fi, f in enumerate(filters):
                if f == "shuffle":
                    piece = np.frombuffer(piece, dtype=np.uint8).reshape(-1, elem).T.tobytes()
                elif f == "deflate":
                    if ci == skip_filter_on:
                        mask |= 1 << fi
                    else:
                        piece = zlib.compress(piece, 6)
                elif f == "fletcher32":
                    piece = piece + struct.pack("<I", fletcher32(piece))
            entries.append((first, len(piece), mask, img.add(piece)))
        if arr.ndim == 2:
            tree = chunk_btree(img, entries, (-(-arr.shape[0] // chunk[0]) * chunk[0], 0)) if entries else UNDEF
            lay = struct.pack("<BBBQIII", 3, 2, 3, tree, chunk[0], chunk[1], elem)
        else:
            tree = chunk_btree(img, entries, -(-len(arr) // chunk) * chunk) if entries else UNDEF
            lay = struct.pack("<BBBQII", 3, 2, 2, tree, chunk, elem)
        if filters:
            extra.append((0x0B, filter_pipeline(filters, elem, 1 if av == 1 else 2)))
    else:
        doff = img.add(raw)
        lay = layout_contiguous(doff, len(raw), 4 if how == "v4" else 3)
    fill = struct.pack("<BBBB", 2, 2, 0, 0)
    if holes:                                                       # a defined fill value: the element 7 (what the reader must return for the holes)
        fv = np.array([7], dtype=arr.dtype.newbyteorder(">" if big else "<")).tobytes()
        fill = struct.pack("<BBBBI", 2, 2, 0, 1, len(fv)) + fv if av == 1 else struct.pack("<BBI", 3, 0x20 | 0x09, len(fv)) + fv
    msgs = [(0x01, dataspace(list(arr.shape), 1 if av == 1 else 2)), (0x03, datatype(arr.dtype, big=big)), (0x05, fill), (0x08, lay)] + extra + [
            (0x0C, attribute("DIMENSION_LIST", vlen_reference_type(), [1], struct.pack("<IQI", 1, UNDEF, 0), av))]
    return hdr(img, msgs, split_after=3 if how == "v4" else None)


def write_old(path):
    start, end, index, data, gatts = sample()
    img = Image()
    sb = img.alloc(96)
    objs = [("examplesDim0", dimension_scale(img, "examplesDim0", 5, 0, header_v1, 1)),
            ("sparseDataDim0", dimension_scale(img, "sparseDataDim0", 12, 1, header_v1, 1)),
            ("sparseStart0", variable(img, "sparseStart0", start, header_v1, 1)),
            ("sparseEnd0", variable(img, "sparseEnd0", end, header_v1, 1)),
            ("sparseIndex0", variable(img, "sparseIndex0", index, header_v1, 1)),
            ("sparseData0", variable(img, "sparseData0", data, header_v1, 1))]
    objs.sort(key=lambda x: x[0])                                 # a symbol table keeps its entries in name order
    heap_data = bytearray(b"\0" * 8)
    name_off = {}
    for n, _ in objs:
        name_off[n] = len(heap_data)
        heap_data += pad8(n.encode() + b"\0")
    hd = img.add(bytes(heap_data))
    heap = img.add(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, hd))
    snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(objs)) + b"".join(struct.pack("<QQII16x", name_off[n], a, 0, 0) for n, a in objs)
    snod += b"\0" * (8 + 32 * 40 - len(snod))                     # a node holds 2 K = 32 entries
    snod_off = img.add(snod)
    tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_off, name_off[objs[-1][0]])
    tree += b"\0" * (24 + 33 * 8 + 32 * 8 - len(tree))
    tree_off = img.add(tree)
    amsgs = [(0x0C, att_text("_NCProperties", "version=2,netcdf=4.7.4,hdf5=1.10.6", 1))]
    for k, v in gatts:
        amsgs.append((0x0C, att_text(k, v, 1) if isinstance(v, str) else att_uint(k, v, 1)))
    root = header_v1(img, [(0x11, struct.pack("<QQ", tree_off, heap))] + amsgs, split_after=4)
    eof = len(img.buf)
    img.put(sb, b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 16, 16, 0) + struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF) +
            struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", tree_off, heap))
    open(path, "wb").write(bytes(img.buf))


def write_new(path):
    start, end, index, data, gatts = sample()
    img = Image()
    sb = img.alloc(48)
    objs = [("examplesDim0", dimension_scale(img, "examplesDim0", 5, 0, header_v2, 3)),
            ("sparseDataDim0", dimension_scale(img, "sparseDataDim0", 12, 1, header_v2, 3)),
            ("sparseStart0", variable(img, "sparseStart0", start, header_v2, 3)),
            ("sparseEnd0", variable(img, "sparseEnd0", end, header_v2, 3, how="compact")),
            ("sparseIndex0", variable(img, "sparseIndex0", index, header_v2, 3, how="v4")),
            ("sparseData0", variable(img, "sparseData0", data, header_v2, 3, big=True))]
    links = fractal_heap(img, [link_message(n, a, i) for i, (n, a) in enumerate(objs)], 512, indirect=False)
    aobjs = [att_text("_NCProperties", "version=2,netcdf=4.8.1,hdf5=1.12.2", 3)]
    for k, v in gatts:
        aobjs.append(att_text(k, v, 3) if isinstance(v, str) else att_uint(k, v, 3))
    atts = fractal_heap(img, aobjs, 160, indirect=True)           # 160-byte blocks: the eight attributes need two direct blocks
    root = header_v2(img, [(0x02, struct.pack("<BBQQQ", 0, 0x01, len(objs), links, UNDEF)), (0x0A, struct.pack("<BB", 0, 0)),
                           (0x15, struct.pack("<BBHQQ", 0, 0x01, len(aobjs), atts, UNDEF))])
    eof = len(img.buf)
    s = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBB", 2, 8, 8, 0) + struct.pack("<QQQQ", 0, UNDEF, eof, root)
    img.put(sb, s + struct.pack("<I", lookup3(s)))
    open(path, "wb").write(bytes(img.buf))


# ---------------------------------------------------------------- any classic netCDF file -> the same content in a netCDF-4 layout
_NCT = {1: np.int8, 2: "S1", 3: np.int16, 4: np.int32, 5: np.float32, 6: np.float64, 7: np.uint8, 8: np.uint16, 9: np.uint32, 10: np.int64, 11: np.uint64}


def read_classic(path):
    """dims [(name, size)], global attributes [(name, value)], variables [(name, dim name, numpy array)] of a CDF-1 / 2 / 5 file whose
    variables are one-dimensional and fixed-size (every file of the DSSTNE schemas)"""
    b = open(path, "rb").read()
    assert b[:3] == b"CDF" and b[3] in (1, 2, 5)
    v = b[3]
    pos = [4]

    def u(n):
        x = int.from_bytes(b[pos[0]:pos[0] + n], "big"); pos[0] += n; return x
    nn = (lambda: u(8)) if v == 5 else (lambda: u(4))

    def name():
        n = nn(); s_ = b[pos[0]:pos[0] + n].decode(); pos[0] += (n + 3) & ~3; return s_

    def atts():
        tag, n = u(4), nn()
        out = []
        for _ in range(n):
            nm, t, ne = name(), u(4), nn()
            if t == 2:
                val = b[pos[0]:pos[0] + ne].decode()
            else:
                val = np.frombuffer(b, dtype=np.dtype(_NCT[t]).newbyteorder(">"), count=ne, offset=pos[0]).astype(_NCT[t])
            pos[0] += (ne * np.dtype(_NCT[t]).itemsize + 3) & ~3
            out.append((nm, val))
        return out
    nn()                                                          # numrecs
    tag, n = u(4), nn()
    dims = [(name(), nn()) for _ in range(n)]
    gatts = atts()
    tag, n = u(4), nn()
    variables = []
    for _ in range(n):
        nm, nd = name(), nn()
        ids = [nn() for _ in range(nd)]
        atts()
        t, vsize, begin = u(4), nn(), (u(4) if v == 1 else u(8))
        assert nd == 1
        cnt = dims[ids[0]][1]
        arr = np.frombuffer(b, dtype=np.dtype(_NCT[t]).newbyteorder(">"), count=cnt, offset=begin).astype(_NCT[t])
        variables.append((nm, dims[ids[0]][0], arr))
    return dims, gatts, variables


def convert(src, dst, flavour, storage=None):
    """flavour "old": superblock v0, version-1 headers, symbol table;  "new": superblock v2, version-2 headers, dense links and attributes.
    storage: None (contiguous) or a function (variable number, name, array) -> keyword arguments of variable() (how / chunk / filters ...)"""
    dims, gatts, variables = read_classic(src)
    img = Image()
    old = flavour == "old"
    hdr, av = (header_v1, 1) if old else (header_v2, 3)
    sb = img.alloc(96 if old else 48)
    objs = [(n, dimension_scale(img, n, size, i, hdr, av)) for i, (n, size) in enumerate(dims)]
    for nm, dn, arr in variables:
        if arr.dtype.kind == "S":
            arr = np.frombuffer(arr.tobytes(), dtype=np.uint8)     # (char variables do not occur in the DSSTNE schemas)
        kw = dict(storage(len(objs), nm, arr)) if storage else {}
        if "shape" in kw:                                          # store a 1-D variable as a 2-D one of the same row-major content
            arr = arr.reshape(kw.pop("shape"))
        objs.append((nm, variable(img, nm, arr, hdr, av, big=(len(objs) % 3 == 0), **kw)))
    amsg = [att_text("_NCProperties", "version=2,netcdf=4.8.1,hdf5=1.12.2", av)]
    for k, val in gatts:
        if isinstance(val, str):
            amsg.append(att_text(k, val, av))
        else:
            amsg.append(attribute(k, datatype(val.dtype), [] if len(val) == 1 else [len(val)], val.astype(val.dtype.newbyteorder("<")).tobytes(), av))
    if old:
        objs.sort(key=lambda x: x[0])
        heap_data = bytearray(b"\0" * 8)
        name_off = {}
        for n, _ in objs:
            name_off[n] = len(heap_data)
            heap_data += pad8(n.encode() + b"\0")
        hd = img.add(bytes(heap_data))
        heap = img.add(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, hd))
        assert len(objs) <= 32
        snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(objs)) + b"".join(struct.pack("<QQII16x", name_off[n], a, 0, 0) for n, a in objs)
        snod += b"\0" * (8 + 32 * 40 - len(snod))
        snod_off = img.add(snod)
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_off, name_off[objs[-1][0]])
        tree += b"\0" * (24 + 33 * 8 + 32 * 8 - len(tree))
        tree_off = img.add(tree)
        root = header_v1(img, [(0x11, struct.pack("<QQ", tree_off, heap))] + [(0x0C, a) for a in amsg], split_after=max(2, len(amsg) // 2))
        eof = len(img.buf)
        img.put(sb, b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 16, 16, 0) + struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF) +
                struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", tree_off, heap))
    else:
        links = fractal_heap(img, [link_message(n, a, i) for i, (n, a) in enumerate(objs)], 512, indirect=True)
        atts = fractal_heap(img, amsg, 512, indirect=True)
        root = header_v2(img, [(0x02, struct.pack("<BBQQQ", 0, 0x01, len(objs), links, UNDEF)), (0x0A, struct.pack("<BB", 0, 0)),
                               (0x15, struct.pack("<BBHQQ", 0, 0x01, len(amsg), atts, UNDEF))])
        eof = len(img.buf)
        sblk = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBB", 2, 8, 8, 0) + struct.pack("<QQQQ", 0, UNDEF, eof, root)
        img.put(sb, sblk + struct.pack("<I", lookup3(sblk)))
    open(dst, "wb").write(bytes(img.buf))


if __name__ == "__main__":
    write_old(os.path.join(HERE, "dataset_nc4_old.nc"))
    write_new(os.path.join(HERE, "dataset_nc4_new.nc"))
    print("wrote", os.path.join(HERE, "dataset_nc4_old.nc"), os.path.join(HERE, "dataset_nc4_new.nc"))
