"""NetCDF classic-format reader / writer of the engine (amazon-dsstne_b200/engine/NetCDF.cpp), on the CPU.

Independent cross-checks with scipy.io.netcdf_file (a separate implementation of CDF-1 / CDF-2):
  * a dataset file in the schema generateNetCDF emits (U/NetCDFhelper.cpp:332-416) written by OUR writer in CDF-2 is read
    back by scipy with identical contents;
  * a file written by scipy (CDF-1 and CDF-2, classic int types) is parsed by OUR reader;
  * CDF-5 (uint / uint64 variables, what the engine writes) round-trips through our own reader, header fields included;
  * netCDF-4 / HDF5 containers (what netcdf-cxx4 writes by default) are read by engine/HDF5.cpp: the two fixtures of
    tests/golden/make_hdf5_fixture.py -- the old-style layout (superblock v0, version-1 object headers, symbol-table root group)
    and the new-style one (superblock v2, version-2 headers, links and attributes in fractal heaps, compact / big-endian
    variables) -- give the same dimensions, attributes and variable contents as the CDF-5 file of the same data set;
  * chunked / compressed netCDF-4 variables, broken containers and non-netCDF files are rejected loudly, not mis-parsed."""
import ctypes as C
import os

import numpy as np
import pytest
from scipy.io import netcdf_file

from helpers import tiny


def describe(lib, path):
    buf = C.create_string_buffer(1 << 16)
    rc = lib.dsb200_netcdf_describe(path.encode(), buf, C.c_size_t(len(buf)))
    return rc, buf.value.decode()


def read_var(lib, path, name):
    n = C.c_uint64()
    assert lib.dsb200_netcdf_read_var(path.encode(), name.encode(), None, C.c_uint64(0), C.byref(n)) == 0
    out = np.zeros(n.value, dtype=np.float64)
    assert lib.dsb200_netcdf_read_var(path.encode(), name.encode(), out.ctypes.data_as(C.c_void_p), C.c_uint64(out.size), C.byref(n)) == 0
    return out


def write_sparse(lib, path, version, name, h, data=None, weight=None, index=None, dtype=0, examples=None):
    uniq = len(h.start)
    lib.dsb200_netcdf_write_sparse.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    return lib.dsb200_netcdf_write_sparse(path.encode(), version, name.encode(), 0, dtype, h.width, examples or uniq, uniq, p(h.start), p(h.end),
                                          p(h.index), p(data), p(weight), p(index))


def test_our_cdf2_writer_is_read_by_scipy(dsb, tmp_path):
    lib = dsb.lib()
    h = tiny(examples=300, width=2048)
    vals = np.random.default_rng(1).uniform(0.5, 5.0, h.nnz).astype(np.float32)
    path = str(tmp_path / "gl_input.nc")
    assert write_sparse(lib, path, 2, "gl_input", h, data=vals, dtype=4) == 0
    with netcdf_file(path, "r", mmap=False) as f:
        assert f.version_byte == 2
        assert f.datasets == 1 and f.name0 == b"gl_input" and f.width0 == 2048 and f.dataType0 == 4 and f.dimensions0 == 1
        assert f.attributes0 == 1                                             # Sparse
        assert f.dimensions["examplesDim0"] == 300 and f.dimensions["sparseDataDim0"] == h.nnz
        np.testing.assert_array_equal(f.variables["sparseStart0"][:], h.start.astype(np.int64))
        np.testing.assert_array_equal(f.variables["sparseEnd0"][:], h.end.astype(np.int64))
        np.testing.assert_array_equal(f.variables["sparseIndex0"][:], h.index.astype(np.int64))
        np.testing.assert_array_equal(f.variables["sparseData0"][:], vals)


@pytest.mark.parametrize("version", [1, 2])
def test_scipy_written_file_is_parsed_by_our_reader(dsb, tmp_path, version):
    lib = dsb.lib()
    h = tiny(examples=64, width=512)
    path = str(tmp_path / f"scipy_v{version}.nc")
    with netcdf_file(path, "w", version=version) as f:
        f.datasets = np.int32(1)
        f.name0 = "gl_output"
        f.attributes0 = np.int32(3)                                           # Sparse + Boolean
        f.kind0 = np.int32(0)
        f.dataType0 = np.int32(0)
        f.dimensions0 = np.int32(1)
        f.width0 = np.int32(512)
        f.createDimension("examplesDim0", 64)
        f.createDimension("sparseDataDim0", h.nnz)
        for name, arr, dim in (("sparseStart0", h.start, "examplesDim0"), ("sparseEnd0", h.end, "examplesDim0"), ("sparseIndex0", h.index, "sparseDataDim0")):
            v = f.createVariable(name, "i4", (dim,))
            v[:] = arr.astype(np.int32)
        w = f.createVariable("dataWeight0", "f4", ("examplesDim0",))
        w[:] = np.linspace(0.5, 1.5, 64, dtype=np.float32)
    rc, text = describe(lib, path)
    assert rc == 0, text
    assert f"(CDF-{version})" in text and "examplesDim0 = 64" in text and 'name0 = "gl_output"' in text and "int sparseIndex0(sparseDataDim0)" in text
    np.testing.assert_array_equal(read_var(lib, path, "sparseStart0"), h.start.astype(np.float64))
    np.testing.assert_array_equal(read_var(lib, path, "sparseEnd0"), h.end.astype(np.float64))
    np.testing.assert_array_equal(read_var(lib, path, "sparseIndex0"), h.index.astype(np.float64))
    np.testing.assert_array_equal(read_var(lib, path, "dataWeight0"), np.linspace(0.5, 1.5, 64, dtype=np.float32).astype(np.float64))


def test_cdf5_round_trip_with_unsigned_types_weights_and_index(dsb, tmp_path):
    lib = dsb.lib()
    h = tiny(examples=100, width=1024)
    rng = np.random.default_rng(2)
    weight = rng.uniform(0.5, 1.5, 100).astype(np.float32)
    index = rng.integers(0, 100, 250).astype(np.uint32)
    vals = rng.integers(0, 255, h.nnz).astype(np.uint8)
    path = str(tmp_path / "indexed.nc")
    assert write_sparse(lib, path, 5, "indexed", h, data=vals, weight=weight, index=index, dtype=8, examples=250) == 0
    rc, text = describe(lib, path)
    assert rc == 0, text
    assert "(CDF-5)" in text and "uniqueExamplesDim0 = 100" in text and "examplesDim0 = 250" in text
    assert "uint sparseStart0(uniqueExamplesDim0)" in text and "ubyte sparseData0(sparseDataDim0)" in text and "uint index0(examplesDim0)" in text
    assert "attributes0 = 193" in text                                        # Sparse | Indexed | Weighted
    np.testing.assert_array_equal(read_var(lib, path, "sparseEnd0"), h.end.astype(np.float64))
    np.testing.assert_array_equal(read_var(lib, path, "sparseData0"), vals.astype(np.float64))
    np.testing.assert_array_equal(read_var(lib, path, "index0"), index.astype(np.float64))
    np.testing.assert_array_equal(read_var(lib, path, "dataWeight0"), weight.astype(np.float64))
    # every variable starts on a 4-byte boundary and the file ends with the last one (classic-format layout rule)
    begins = [int(l.split("begin=")[1].split()[0]) for l in text.splitlines() if "begin=" in l]
    sizes = [int(l.split("vsize=")[1].split()[0]) for l in text.splitlines() if "vsize=" in l]
    assert all(b % 4 == 0 for b in begins) and begins == sorted(begins)
    assert os.path.getsize(path) == begins[-1] + sizes[-1]


def _fixture_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_hdf5_fixture", os.path.join(os.path.dirname(__file__), "golden", "make_hdf5_fixture.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("flavour", ["old", "new"])
def test_netcdf4_hdf5_fixture_reads_like_the_cdf5_file_of_the_same_data_set(dsb, tmp_path, flavour):
    lib = dsb.lib()
    m = _fixture_module()
    golden = os.path.join(os.path.dirname(__file__), "golden", f"dataset_nc4_{flavour}.nc")
    fresh = str(tmp_path / f"nc4_{flavour}.nc")
    (m.write_old if flavour == "old" else m.write_new)(fresh)
    assert open(fresh, "rb").read() == open(golden, "rb").read()              # the committed fixture is what the committed script writes
    start, end, index, data, gatts = m.sample()
    rc, text = describe(lib, golden)
    assert rc == 0, text
    assert "netCDF-4 / HDF5" in text and "examplesDim0 = 5" in text and "sparseDataDim0 = 12" in text
    for hidden in ("_NCProperties", "DIMENSION_LIST", "_Netcdf4Dimid", "CLASS", "NAME"):
        assert hidden not in text                                               # netCDF-4 bookkeeping is not user data
    for k, v in gatts:
        assert (f'{k} = "{v}"' if isinstance(v, str) else f"{k} = {v} (uint)") in text
    assert "uint sparseStart0(examplesDim0)" in text and "float sparseData0(sparseDataDim0)" in text
    np.testing.assert_array_equal(read_var(lib, golden, "sparseStart0"), start)
    np.testing.assert_array_equal(read_var(lib, golden, "sparseEnd0"), end)
    np.testing.assert_array_equal(read_var(lib, golden, "sparseIndex0"), index)
    np.testing.assert_array_equal(read_var(lib, golden, "sparseData0"), data)
    # the same data set written by our CDF-5 writer reads back identically
    class H: pass
    h = H(); h.start, h.end, h.index, h.width = start.astype(np.uint64), end.astype(np.uint64), index, 256
    cdf5 = str(tmp_path / "same.nc")
    assert write_sparse(lib, cdf5, 5, "gl_input", h, data=data, dtype=4) == 0
    for v in ("sparseStart0", "sparseEnd0", "sparseIndex0", "sparseData0"):
        np.testing.assert_array_equal(read_var(lib, golden, v), read_var(lib, cdf5, v))


@pytest.mark.parametrize("flavour", ["old", "new"])
def test_netcdf4_container_with_a_network_sized_header(dsb, tmp_path, flavour):
    """a header of the size of a network checkpoint (E/NNNetwork.cpp:1936-1970 writes ~30 global attributes plus ~25 per layer and
    weight): 160 attributes of mixed types and a dozen variables, written by scipy as CDF-2 and re-laid-out as netCDF-4 by the
    fixture script -- the dense attribute heap then spans several rows of its doubling table; both files must read the same"""
    lib = dsb.lib()
    m = _fixture_module()
    rng = np.random.default_rng(7)
    src, dst = str(tmp_path / "net.nc"), str(tmp_path / f"net_{flavour}.nc")
    arrays = {}
    with netcdf_file(src, "w", version=2) as f:
        for i in range(160):
            k = i % 4
            setattr(f, f"att{i:03d}_{'ifsx'[k]}", [np.int32(rng.integers(-1000, 1000)), np.float32(rng.standard_normal()), f"layer{i}_name with spaces",
                                                  np.array(rng.integers(0, 9, 5), dtype=np.int32)][k])
        for j in range(12):
            n = int(rng.integers(1, 400))
            f.createDimension(f"dim{j}", n)
            dt = ["f", "i", "d"][j % 3]
            v = f.createVariable(f"var{j}", dt, (f"dim{j}",))
            arrays[f"var{j}"] = (rng.standard_normal(n) * 100).astype({"f": np.float32, "i": np.int32, "d": np.float64}[dt])
            v[:] = arrays[f"var{j}"]
    m.convert(src, dst, flavour)
    rc0, want = describe(lib, src)
    rc1, got = describe(lib, dst)
    assert rc0 == 0 and rc1 == 0, got

    def body(text):                                              # attributes, dimensions, variable types and names; not offsets / container names
        keep = []
        for line in text.splitlines()[1:]:
            keep.append(line.split(" begin=")[0])
        return sorted(keep)
    assert body(got) == body(want)
    for name, arr in arrays.items():
        np.testing.assert_array_equal(read_var(lib, dst, name), arr.astype(np.float64))


_STORAGE = {
    "chunked": lambda i, nm, a: dict(how="chunked", chunk=37),
    "deflate": lambda i, nm, a: dict(how="chunked", chunk=64, filters=("deflate",)),
    "shuffle+deflate": lambda i, nm, a: dict(how="chunked", chunk=50, filters=("shuffle", "deflate")),
    "shuffle+deflate+fletcher32": lambda i, nm, a: dict(how="chunked", chunk=128, filters=("shuffle", "deflate", "fletcher32"), skip_filter_on=1),
    "fletcher32+deflate": lambda i, nm, a: dict(how="chunked", chunk=41, filters=("fletcher32", "deflate")),
    "two-level index": lambda i, nm, a: dict(how="chunked", chunk=3, filters=("shuffle", "deflate") if i % 2 else ()),
    "mixed": lambda i, nm, a: [dict(), dict(how="compact"), dict(how="chunked", chunk=1000, filters=("deflate",)), dict(how="chunked", chunk=1)][i % 4],
}


@pytest.mark.parametrize("flavour", ["old", "new"])
@pytest.mark.parametrize("storage", sorted(_STORAGE))
def test_netcdf4_chunked_and_compressed_variables(dsb, tmp_path, flavour, storage):
    """what `nccopy -d N -s` or a writer with an unlimited examples dimension produces: variables cut into chunks behind a version-1
    B-tree (one and two levels: 2 K = 64 children per node), stored through shuffle / deflate / fletcher32 in either order, one chunk
    with its deflate step skipped (mask bit), chunks that do not divide the length; all seven element types of the DSSTNE schemas"""
    lib = dsb.lib()
    m = _fixture_module()
    rng = np.random.default_rng(11)
    src, dst = str(tmp_path / "d.nc"), str(tmp_path / f"d_{flavour}.nc")
    arrays = {}
    with netcdf_file(src, "w", version=2) as f:
        f.datasets = np.int32(1)
        for j, dt in enumerate(["f", "i", "d", "b", "h", "i", "f", "d"]):
            n = [611, 257, 1, 300, 129, 64, 5000, 200][j]
            f.createDimension(f"dim{j}", n)
            v = f.createVariable(f"var{j}", dt, (f"dim{j}",))
            np_dt = {"f": np.float32, "i": np.int32, "d": np.float64, "b": np.int8, "h": np.int16}[dt]
            arrays[f"var{j}"] = (rng.integers(-5, 5, n) * 3).astype(np_dt) if j % 2 else (rng.standard_normal(n) * 50).astype(np_dt)
            v[:] = arrays[f"var{j}"]
    m.convert(src, dst, flavour, storage=_STORAGE[storage])
    rc, text = describe(lib, dst)
    assert rc == 0, text
    for name, arr in arrays.items():
        np.testing.assert_array_equal(read_var(lib, dst, name), arr.astype(np.float64), err_msg=f"{name} ({storage}, {flavour})")


@pytest.mark.parametrize("flavour", ["old", "new"])
def test_netcdf4_two_dimensional_chunks(dsb, tmp_path, flavour):
    """not a DSSTNE layout, but what a generic tool may hand over: a 2-D variable in (rows, columns) tiles that overhang both edges;
    read_var returns it row-major"""
    lib = dsb.lib()
    m = _fixture_module()
    src, dst = str(tmp_path / "t.nc"), str(tmp_path / f"t_{flavour}.nc")
    a = (np.random.default_rng(5).standard_normal(23 * 17) * 9).astype(np.float32)
    with netcdf_file(src, "w", version=2) as f:
        f.createDimension("n", a.size)
        f.createVariable("v", "f", ("n",))[:] = a
        f.createVariable("w", "f", ("n",))[:] = -a
    m.convert(src, dst, flavour, storage=lambda i, nm, arr: dict(how="chunked", shape=(23, 17), chunk=(5, 4) if nm == "v" else (23, 17),
                                                               filters=("shuffle", "deflate") if nm == "v" else (), holes=(3,) if nm == "v" else ()))
    want = a.astype(np.float64).reshape(23, 17).copy()
    want[0:5, 12:16] = 7                                                      # chunk number 3 of the first row of tiles
    np.testing.assert_array_equal(read_var(lib, dst, "v").reshape(23, 17), want)
    np.testing.assert_array_equal(read_var(lib, dst, "w"), -a.astype(np.float64))


@pytest.mark.parametrize("flavour", ["old", "new"])
def test_netcdf4_never_written_chunks_read_as_the_fill_value(dsb, tmp_path, flavour):
    lib = dsb.lib()
    m = _fixture_module()
    src, dst = str(tmp_path / "h.nc"), str(tmp_path / f"h_{flavour}.nc")
    a = np.arange(100, dtype=np.int32)
    with netcdf_file(src, "w", version=2) as f:
        f.createDimension("n", 100)
        v = f.createVariable("v", "i", ("n",))
        v[:] = a
        w = f.createVariable("w", "f", ("n",))
        w[:] = a.astype(np.float32)
    m.convert(src, dst, flavour, storage=lambda i, nm, arr: dict(how="chunked", chunk=16, holes=(1, 6) if nm == "v" else tuple(range(7)), filters=("deflate",)))
    want = a.astype(np.float64).copy()
    want[16:32] = 7
    want[96:] = 7
    np.testing.assert_array_equal(read_var(lib, dst, "v"), want)
    np.testing.assert_array_equal(read_var(lib, dst, "w"), np.full(100, 7.0))     # no chunk written at all: no index either


def test_unreadable_containers_are_rejected_loudly(dsb, tmp_path):
    lib = dsb.lib()
    lib.dsb200_engine_last_error.restype = C.c_char_p
    p1 = tmp_path / "nc4.nc"
    p1.write_bytes(b"\x89HDF\r\n\x1a\n" + bytes(64))                          # a signature and nothing behind it
    rc, _ = describe(lib, str(p1))
    assert rc != 0 and b"HDF5" in lib.dsb200_engine_last_error()
    # chunk indexes of HDF5's "latest format" (layout message version 4, class chunked) and foreign filters: named, with the way out
    m = _fixture_module()
    import struct
    orig = m.layout_contiguous
    try:
        m.layout_contiguous = lambda addr, size, version=3: struct.pack("<BBBBBQ", 4, 2, 0, 1, 4, addr) if addr != m.UNDEF else orig(addr, size, version)
        p4 = str(tmp_path / "v4index.nc")
        m.write_old(p4)
    finally:
        m.layout_contiguous = orig
    rc, _ = describe(lib, p4)
    err = lib.dsb200_engine_last_error()
    assert rc != 0 and b"version-4 chunk index" in err and b"nccopy" in err
    origp = m.filter_pipeline
    try:
        m.filter_pipeline = lambda filters, elem, version: origp(filters, elem, version).replace(struct.pack("<H", 1), struct.pack("<H", 4), 1) if version == 2 else origp(filters, elem, version)
        src, p5 = str(tmp_path / "z.nc"), str(tmp_path / "szip.nc")
        with netcdf_file(src, "w", version=2) as f:
            f.createDimension("n", 10)
            f.createVariable("v", "i", ("n",))[:] = np.arange(10, dtype=np.int32)
        m.convert(src, p5, "new", storage=lambda i, nm, a: dict(how="chunked", chunk=4, filters=("deflate",)))
    finally:
        m.filter_pipeline = origp
    rc, _ = describe(lib, p5)
    err = lib.dsb200_engine_last_error()
    assert rc != 0 and b"filter 4;" in err and b"nccopy" in err, err
    # a chunk that does not inflate: the error names the variable and the offset
    p6 = str(tmp_path / "corrupt.nc")
    m.convert(src, p6, "old", storage=lambda i, nm, a: dict(how="chunked", chunk=4, filters=("deflate",)))
    blob = bytearray(open(p6, "rb").read())
    at = blob.index(b"\x78\x9c")                                            # first zlib stream
    blob[at + 2:at + 6] = b"\xff\xff\xff\xff"
    open(p6, "wb").write(bytes(blob))
    assert describe(lib, p6)[0] == 0                                         # the header is fine ...
    out = np.zeros(10)
    n = C.c_uint64()
    rc = lib.dsb200_netcdf_read_var(p6.encode(), b"v", out.ctypes.data_as(C.c_void_p), C.c_uint64(10), C.byref(n))
    err = lib.dsb200_engine_last_error()
    assert rc != 0 and b"variable v" in err and b"does not inflate" in err, err       # ... the data is not
    p2 = tmp_path / "junk.nc"
    p2.write_bytes(b"hello world, not a netcdf file")
    rc, _ = describe(lib, str(p2))
    assert rc != 0 and b"bad magic" in lib.dsb200_engine_last_error()
    p3 = tmp_path / "short.nc"
    p3.write_bytes(b"CDF\x05\x00\x00")
    rc, _ = describe(lib, str(p3))
    assert rc != 0 and b"truncated" in lib.dsb200_engine_last_error()


_FUZZ = r"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import dsstne_b200
lib = dsstne_b200.lib()
blob = open(sys.argv[2], "rb").read()
tmp = sys.argv[3]
rng = np.random.default_rng(3)
buf = C.create_string_buffer(1 << 16)
out = np.zeros(4096, dtype=np.float64)
cases = [blob[:n] for n in range(0, len(blob), 5)]
for i in range(400):
    b = bytearray(blob)
    for _ in range(1 + i % 3):
        b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
    cases.append(bytes(b))
bad = 0
for b in cases:
    open(tmp, "wb").write(b)
    bad += lib.dsb200_netcdf_describe(tmp.encode(), buf, C.c_size_t(len(buf))) != 0
    for name in (b"sparseStart0", b"sparseIndex0", b"sparseData0", b"v"):
        n = C.c_uint64()
        lib.dsb200_netcdf_read_var(tmp.encode(), name, out.ctypes.data_as(C.c_void_p), C.c_uint64(out.size), C.byref(n))
print("survived", len(cases), "rejected", bad)
"""


@pytest.mark.parametrize("which", ["old", "new", "chunked"])
def test_damaged_netcdf4_containers_never_crash_the_reader(dsb, tmp_path, which):
    """every prefix (step 5) of a container and 400 copies with one to three bytes overwritten: the reader either reads the file or
    returns an error -- no crash, no hang, no unbounded allocation (run in a child process so that a crash fails only this test)"""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    if which == "chunked":
        m = _fixture_module()
        src, path = str(tmp_path / "c.nc"), str(tmp_path / "c4.nc")
        with netcdf_file(src, "w", version=2) as f:
            f.createDimension("n", 700)
            f.createVariable("v", "f", ("n",))[:] = np.arange(700, dtype=np.float32)
            f.createVariable("sparseIndex0", "i", ("n",))[:] = np.arange(700, dtype=np.int32)
        m.convert(src, path, "old", storage=lambda i, nm, a: dict(how="chunked", chunk=3 if nm == "v" else 64, filters=("shuffle", "deflate", "fletcher32")))
    else:
        path = os.path.join(here, "golden", f"dataset_nc4_{which}.nc")
    script = tmp_path / "fuzz.py"
    script.write_text(_FUZZ)
    r = subprocess.run([sys.executable, str(script), os.path.dirname(here), path, str(tmp_path / "damaged.nc")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "survived" in r.stdout, (r.returncode, r.stdout[-300:], r.stderr[-600:])
    rejected = int(r.stdout.split()[-1])
    assert rejected > 100                                                     # the truncated ones at the very least
