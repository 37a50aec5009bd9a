// NetCDF.cpp -- netCDF classic (CDF-1 / CDF-2 / CDF-5) header parser, variable reader and writer.  See NetCDF.h.
//
// File grammar (netCDF "classic format specification"):
//   header   = magic numrecs dim_list gatt_list var_list
//   magic    = 'C' 'D' 'F' version(1|2|5)
//   NON_NEG  = 4 bytes big-endian (CDF-1/2) or 8 bytes (CDF-5): counts, lengths, numrecs, vsize, dimids
//   OFFSET   = 4 bytes (CDF-1) or 8 bytes (CDF-2/5): var.begin
//   list     = ABSENT (tag 0, count 0) | tag(NC_DIMENSION 10 | NC_VARIABLE 11 | NC_ATTRIBUTE 12) count entries
//   name     = NON_NEG length, bytes, zero padding to a multiple of 4
//   att      = name nc_type(4 bytes) NON_NEG nelems, values, padding to 4
//   var      = name NON_NEG ndims, dimids, vatt_list, nc_type, NON_NEG vsize, OFFSET begin
//   data     = fixed-size variables at their `begin`, big-endian, each padded to 4 bytes
#include "NetCDF.h"

#include <cstdio>
#include <cstring>
#include <sstream>

namespace nc {

size_t type_size(Type t)
{
    switch (t) {
    case NC_BYTE: case NC_CHAR: case NC_UBYTE: return 1;
    case NC_SHORT: case NC_USHORT: return 2;
    case NC_INT: case NC_FLOAT: case NC_UINT: return 4;
    case NC_DOUBLE: case NC_INT64: case NC_UINT64: return 8;
    }
    throw Error("netCDF: unknown type code " + std::to_string((int)t));
}

const char* type_name(Type t)
{
    static const char* n[] = {"?", "byte", "char", "short", "int", "float", "double", "ubyte", "ushort", "uint", "int64", "uint64"};
    return (t >= 1 && t <= 11) ? n[t] : "?";
}

// big-endian file bytes <-> host values
static void swap_elems(uint8_t* p, size_t n, size_t sz)
{
    if (sz == 1) return;
    const uint16_t probe = 1;
    if (*reinterpret_cast<const uint8_t*>(&probe) == 0) return;           // big-endian host: file order already
    for (size_t i = 0; i < n; i++, p += sz)
        for (size_t a = 0, b = sz - 1; a < b; a++, b--) { uint8_t t = p[a]; p[a] = p[b]; p[b] = t; }
}

template <typename T> static T host_value(const uint8_t* p, Type t)
{
    switch (t) {
    case NC_BYTE:   { int8_t v;   memcpy(&v, p, 1); return (T)v; }
    case NC_CHAR:
    case NC_UBYTE:  { uint8_t v;  memcpy(&v, p, 1); return (T)v; }
    case NC_SHORT:  { int16_t v;  memcpy(&v, p, 2); return (T)v; }
    case NC_USHORT: { uint16_t v; memcpy(&v, p, 2); return (T)v; }
    case NC_INT:    { int32_t v;  memcpy(&v, p, 4); return (T)v; }
    case NC_UINT:   { uint32_t v; memcpy(&v, p, 4); return (T)v; }
    case NC_FLOAT:  { float v;    memcpy(&v, p, 4); return (T)v; }
    case NC_DOUBLE: { double v;   memcpy(&v, p, 8); return (T)v; }
    case NC_INT64:  { int64_t v;  memcpy(&v, p, 8); return (T)v; }
    case NC_UINT64: { uint64_t v; memcpy(&v, p, 8); return (T)v; }
    }
    throw Error("netCDF: unknown type code");
}

std::string Att::as_string() const
{
    if (type != NC_CHAR) throw Error("netCDF: attribute " + name + " is not text");
    std::string s(data.begin(), data.end());
    while (!s.empty() && s.back() == '\0') s.pop_back();
    return s;
}
double Att::as_double(size_t i) const
{
    if (type == NC_CHAR || i >= nelems) throw Error("netCDF: attribute " + name + " has no numeric element " + std::to_string(i));
    return host_value<double>(data.data() + i * type_size(type), type);
}
int64_t Att::as_int(size_t i) const
{
    if (type == NC_CHAR || i >= nelems) throw Error("netCDF: attribute " + name + " has no numeric element " + std::to_string(i));
    return host_value<int64_t>(data.data() + i * type_size(type), type);
}

namespace {

struct Cursor {
    FILE* f;
    int version;
    const std::string& fname;
    void bytes(void* dst, size_t n)
    {
        if (n && fread(dst, 1, n, f) != n) throw Error("netCDF: truncated header in " + fname);
    }
    uint32_t u32() { uint8_t b[4]; bytes(b, 4); return ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3]; }
    uint64_t u64() { const uint64_t hi = u32(); return (hi << 32) | u32(); }
    uint64_t non_neg() { return version == 5 ? u64() : u32(); }
    uint64_t offset() { return version == 1 ? u32() : u64(); }
    void skip_pad(uint64_t n) { const uint64_t pad = (4 - (n & 3)) & 3; uint8_t b[4]; bytes(b, pad); }
    std::string name()
    {
        const uint64_t n = non_neg();
        if (n > (1u << 20)) throw Error("netCDF: implausible name length in " + fname);
        std::string s(n, '\0');
        bytes(&s[0], n);
        skip_pad(n);
        return s;
    }
    Type type()
    {
        const uint32_t t = u32();
        if (t < 1 || t > 11 || (version != 5 && t > 6)) throw Error("netCDF: invalid type code " + std::to_string(t) + " in " + fname);
        return (Type)t;
    }
    void att_list(std::vector<Att>& out)
    {
        const uint32_t tag = u32();
        const uint64_t n = non_neg();
        if (tag == 0 && n == 0) return;
        if (tag != 12) throw Error("netCDF: expected an attribute list in " + fname);
        for (uint64_t i = 0; i < n; i++) {
            Att a;
            a.name = name();
            a.type = type();
            a.nelems = non_neg();
            const uint64_t nbytes = a.nelems * type_size(a.type);
            if (nbytes > (1ull << 32)) throw Error("netCDF: implausible attribute size in " + fname);
            a.data.resize(nbytes);
            bytes(a.data.data(), nbytes);
            skip_pad(nbytes);
            swap_elems(a.data.data(), a.nelems, type_size(a.type));
            out.push_back(a);
        }
    }
};

}  // namespace

File::File(const std::string& fname) : _fname(fname), _version(0), _numrecs(0)
{
    FILE* f = fopen(fname.c_str(), "rb");
    if (!f) throw Error("netCDF: cannot open " + fname);
    struct Closer { FILE* f; ~Closer() { fclose(f); } } closer{f};
    uint8_t magic[8] = {0};
    if (fread(magic, 1, 4, f) != 4) throw Error("netCDF: " + fname + " is too short to be a netCDF file");
    if (magic[0] == 0x89 && magic[1] == 'H' && magic[2] == 'D' && magic[3] == 'F') {
        _version = 4;                                                          // netCDF-4: an HDF5 container (HDF5.cpp)
        hdf5::parse(f, fname, _dims, _atts, _vars);
        return;
    }
    if (magic[0] != 'C' || magic[1] != 'D' || magic[2] != 'F' || (magic[3] != 1 && magic[3] != 2 && magic[3] != 5))
        throw Error("netCDF: " + fname + " is not a netCDF classic file (bad magic)");
    _version = magic[3];
    Cursor c{f, _version, _fname};
    _numrecs = c.non_neg();
    {   // dim_list
        const uint32_t tag = c.u32();
        const uint64_t n = c.non_neg();
        if (!(tag == 0 && n == 0)) {
            if (tag != 10) throw Error("netCDF: expected a dimension list in " + fname);
            for (uint64_t i = 0; i < n; i++) {
                Dim d;
                d.name = c.name();
                d.size = c.non_neg();
                _dims.push_back(d);
            }
        }
    }
    c.att_list(_atts);
    {   // var_list
        const uint32_t tag = c.u32();
        const uint64_t n = c.non_neg();
        if (!(tag == 0 && n == 0)) {
            if (tag != 11) throw Error("netCDF: expected a variable list in " + fname);
            for (uint64_t i = 0; i < n; i++) {
                Var v;
                v.name = c.name();
                const uint64_t ndims = c.non_neg();
                if (ndims > 1024) throw Error("netCDF: implausible rank in " + fname);
                v.nelems = 1;
                for (uint64_t k = 0; k < ndims; k++) {
                    const uint64_t id = c.non_neg();
                    if (id >= _dims.size()) throw Error("netCDF: variable " + v.name + " refers to an unknown dimension in " + fname);
                    if (_dims[id].size == 0) throw Error("netCDF: variable " + v.name + " uses the record dimension, which DSSTNE files never do (" + fname + ")");
                    v.dimids.push_back((uint32_t)id);
                    v.nelems *= _dims[id].size;
                }
                c.att_list(v.atts);
                v.type = c.type();
                v.vsize = c.non_neg();
                v.begin = c.offset();
                _vars.push_back(v);
            }
        }
    }
}

const Att* File::att(const std::string& name) const
{
    for (const Att& a : _atts) if (a.name == name) return &a;
    return nullptr;
}
const Dim* File::dim(const std::string& name) const
{
    for (const Dim& d : _dims) if (d.name == name) return &d;
    return nullptr;
}
const Var* File::var(const std::string& name) const
{
    for (const Var& v : _vars) if (v.name == name) return &v;
    return nullptr;
}

void File::read_raw(const Var& v, std::vector<uint8_t>& bytes) const
{
    FILE* f = fopen(_fname.c_str(), "rb");
    if (!f) throw Error("netCDF: cannot reopen " + _fname);
    struct Closer { FILE* f; ~Closer() { fclose(f); } } closer{f};
    const uint64_t n = v.nelems * type_size(v.type);
    if (v.hasInline) {
        if (v.inlineData.size() < n) throw Error("netCDF: variable " + v.name + " is truncated in " + _fname);
        bytes.assign(v.inlineData.begin(), v.inlineData.begin() + n);
        return;
    }
    if (v.chunked) {
        hdf5::read_chunked(f, _fname, v, bytes);
        if (bytes.size() != n) throw Error("netCDF: variable " + v.name + " has " + std::to_string(bytes.size()) + " bytes of chunks for " + std::to_string(n) + " in " + _fname);
        return;
    }
    if (fseeko(f, 0, SEEK_END) == 0) {
        const uint64_t size = (uint64_t)ftello(f);
        if (v.begin > size || n > size - v.begin) throw Error("netCDF: variable " + v.name + " is truncated in " + _fname);
    }
    bytes.resize(n);
#if defined(_WIN32)
    if (_fseeki64(f, (long long)v.begin, SEEK_SET) != 0)
#else
    if (fseeko(f, (off_t)v.begin, SEEK_SET) != 0)
#endif
        throw Error("netCDF: cannot seek to variable " + v.name + " in " + _fname);
    if (n && fread(bytes.data(), 1, n, f) != n) throw Error("netCDF: variable " + v.name + " is truncated in " + _fname);
}

template <typename T> void File::read(const Var& v, std::vector<T>& out) const
{
    std::vector<uint8_t> raw;
    read_raw(v, raw);
    const size_t sz = type_size(v.type);
    if (!v.littleEndian) swap_elems(raw.data(), v.nelems, sz);                // classic files are big-endian
    else {
        const uint16_t probe = 1;
        if (*reinterpret_cast<const uint8_t*>(&probe) == 0 && sz > 1)         // little-endian data on a big-endian host
            for (uint64_t i = 0; i < v.nelems; i++) for (size_t a = 0, b = sz - 1; a < b; a++, b--) std::swap(raw[i * sz + a], raw[i * sz + b]);
    }
    out.resize(v.nelems);
    for (uint64_t i = 0; i < v.nelems; i++) out[i] = host_value<T>(raw.data() + i * sz, v.type);
}
template void File::read<uint8_t>(const Var&, std::vector<uint8_t>&) const;
template void File::read<int8_t>(const Var&, std::vector<int8_t>&) const;
template void File::read<char>(const Var&, std::vector<char>&) const;
template void File::read<int32_t>(const Var&, std::vector<int32_t>&) const;
template void File::read<uint32_t>(const Var&, std::vector<uint32_t>&) const;
template void File::read<int64_t>(const Var&, std::vector<int64_t>&) const;
template void File::read<uint64_t>(const Var&, std::vector<uint64_t>&) const;
template void File::read<float>(const Var&, std::vector<float>&) const;
template void File::read<double>(const Var&, std::vector<double>&) const;

std::string File::describe() const
{
    std::ostringstream o;
    o << "netcdf " << _fname << (_version == 4 ? " (netCDF-4 / HDF5" : " (CDF-") << (_version == 4 ? std::string() : std::to_string(_version)) << ")\n" << "dimensions:\n";
    for (const Dim& d : _dims) o << "\t" << d.name << " = " << d.size << "\n";
    o << "variables:\n";
    for (const Var& v : _vars) {
        o << "\t" << type_name(v.type) << " " << v.name << "(";
        for (size_t i = 0; i < v.dimids.size(); i++) o << (i ? ", " : "") << _dims[v.dimids[i]].name;
        o << ") begin=" << v.begin << " vsize=" << v.vsize << "\n";
    }
    o << "global attributes:\n";
    for (const Att& a : _atts) {
        o << "\t" << a.name << " = ";
        if (a.type == NC_CHAR) o << "\"" << a.as_string() << "\"";
        else for (uint64_t i = 0; i < a.nelems; i++) o << (i ? ", " : "") << a.as_double(i);
        o << " (" << type_name(a.type) << ")\n";
    }
    return o.str();
}

// ---------------------------------------------------------------- writer

Writer::Writer(int version) : _version(version)
{
    if (version != 2 && version != 5) throw Error("netCDF writer: format must be 2 (CDF-2) or 5 (CDF-5)");
}

void Writer::add_dim(const std::string& name, uint64_t size)
{
    if (size == 0) throw Error("netCDF writer: dimension " + name + " has size 0 (the record dimension is not supported)");
    if (_version != 5 && size > 0x7fffffffu) throw Error("netCDF writer: dimension " + name + " needs CDF-5");
    _dims.push_back(Dim{name, size});
}

void Writer::put_att(const std::string& name, const std::string& value)
{
    Att a;
    a.name = name; a.type = NC_CHAR; a.nelems = value.size();
    a.data.assign(value.begin(), value.end());
    _atts.push_back(a);
}

void Writer::put_att(const std::string& name, Type type, double value)
{
    if (type == NC_INT64 || type == NC_UINT64) { put_att_u64(name, type, (uint64_t)value); return; }
    Att a;
    a.name = name; a.nelems = 1;
    if (_version != 5 && type > NC_DOUBLE) type = (type == NC_UBYTE) ? NC_BYTE : (type == NC_USHORT) ? NC_SHORT : NC_INT;   // classic has no unsigned types
    a.type = type;
    a.data.resize(type_size(type));
    switch (type) {
    case NC_BYTE:   { int8_t v = (int8_t)value;     memcpy(a.data.data(), &v, 1); break; }
    case NC_UBYTE:  { uint8_t v = (uint8_t)value;   memcpy(a.data.data(), &v, 1); break; }
    case NC_SHORT:  { int16_t v = (int16_t)value;   memcpy(a.data.data(), &v, 2); break; }
    case NC_USHORT: { uint16_t v = (uint16_t)value; memcpy(a.data.data(), &v, 2); break; }
    case NC_INT:    { int32_t v = (int32_t)value;   memcpy(a.data.data(), &v, 4); break; }
    case NC_UINT:   { uint32_t v = (uint32_t)value; memcpy(a.data.data(), &v, 4); break; }
    case NC_FLOAT:  { float v = (float)value;       memcpy(a.data.data(), &v, 4); break; }
    case NC_DOUBLE: { double v = value;             memcpy(a.data.data(), &v, 8); break; }
    default: throw Error("netCDF writer: bad attribute type");
    }
    _atts.push_back(a);
}

void Writer::put_att_u64(const std::string& name, Type type, uint64_t value)
{
    if (_version != 5) {                              // classic: the widest integer is a 32-bit int; fall back to double beyond it
        if (value <= 0x7fffffffull) { put_att(name, NC_INT, (double)value); return; }
        put_att(name, NC_DOUBLE, (double)value);
        return;
    }
    Att a;
    a.name = name; a.type = type; a.nelems = 1;
    a.data.resize(8);
    memcpy(a.data.data(), &value, 8);
    _atts.push_back(a);
}

void Writer::add_var(const std::string& name, Type type, const std::string& dim, const void* data)
{
    if (_version != 5 && type > NC_DOUBLE) throw Error("netCDF writer: variable " + name + " of type " + type_name(type) + " needs CDF-5");
    for (uint32_t i = 0; i < _dims.size(); i++)
        if (_dims[i].name == dim) { _vars.push_back(V{name, type, i, data}); return; }
    throw Error("netCDF writer: variable " + name + " uses undefined dimension " + dim);
}

namespace {
struct Out {
    std::vector<uint8_t> b;
    int version;
    void u32(uint32_t v) { b.push_back(v >> 24); b.push_back(v >> 16); b.push_back(v >> 8); b.push_back(v); }
    void u64(uint64_t v) { u32((uint32_t)(v >> 32)); u32((uint32_t)v); }
    void non_neg(uint64_t v) { if (version == 5) u64(v); else u32((uint32_t)v); }
    void pad() { while (b.size() & 3) b.push_back(0); }
    void name(const std::string& s) { non_neg(s.size()); b.insert(b.end(), s.begin(), s.end()); pad(); }
    void att(const Att& a)
    {
        name(a.name);
        u32((uint32_t)a.type);
        non_neg(a.nelems);
        std::vector<uint8_t> d = a.data;
        swap_elems(d.data(), a.nelems, type_size(a.type));
        b.insert(b.end(), d.begin(), d.end());
        pad();
    }
};
}  // namespace

void Writer::write(const std::string& fname) const
{
    // pass 1 sizes the header (begin offsets depend on it), pass 2 emits it
    uint64_t headerBytes = 0;
    std::vector<uint64_t> begin(_vars.size()), vsize(_vars.size());
    Out h;
    for (int pass = 0; pass < 2; pass++) {
        h = Out();
        h.version = _version;
        h.b.push_back('C'); h.b.push_back('D'); h.b.push_back('F'); h.b.push_back((uint8_t)_version);
        h.non_neg(0);                                                   // numrecs
        if (_dims.empty()) { h.u32(0); h.non_neg(0); }
        else {
            h.u32(10); h.non_neg(_dims.size());
            for (const Dim& d : _dims) { h.name(d.name); h.non_neg(d.size); }
        }
        if (_atts.empty()) { h.u32(0); h.non_neg(0); }
        else {
            h.u32(12); h.non_neg(_atts.size());
            for (const Att& a : _atts) h.att(a);
        }
        if (_vars.empty()) { h.u32(0); h.non_neg(0); }
        else {
            h.u32(11); h.non_neg(_vars.size());
            uint64_t off = headerBytes;
            for (size_t i = 0; i < _vars.size(); i++) {
                const V& v = _vars[i];
                const uint64_t bytes = _dims[v.dimid].size * type_size(v.type);
                vsize[i] = (bytes + 3) & ~3ull;
                begin[i] = off;
                off += vsize[i];
                h.name(v.name);
                h.non_neg(1);
                h.non_neg(v.dimid);
                h.u32(0); h.non_neg(0);                                 // no per-variable attributes
                h.u32((uint32_t)v.type);
                if (_version != 5 && vsize[i] > 0xffffffffull) throw Error("netCDF writer: variable " + v.name + " needs CDF-5");
                h.non_neg(vsize[i]);
                h.u64(begin[i]);                                        // OFFSET is 64-bit in CDF-2 and CDF-5
            }
        }
        headerBytes = h.b.size();
    }
    FILE* f = fopen(fname.c_str(), "wb");
    if (!f) throw Error("netCDF writer: cannot create " + fname);
    struct Closer { FILE* f; ~Closer() { fclose(f); } } closer{f};
    if (fwrite(h.b.data(), 1, h.b.size(), f) != h.b.size()) throw Error("netCDF writer: short write to " + fname);
    std::vector<uint8_t> buf;
    for (size_t i = 0; i < _vars.size(); i++) {
        const V& v = _vars[i];
        const size_t sz = type_size(v.type);
        const uint64_t n = _dims[v.dimid].size;
        const uint64_t chunk = 1u << 20;                                // elements per pass: bounded scratch for multi-GB variables
        for (uint64_t o = 0; o < n; o += chunk) {
            const uint64_t m = (n - o < chunk) ? n - o : chunk;
            buf.assign(static_cast<const uint8_t*>(v.data) + o * sz, static_cast<const uint8_t*>(v.data) + (o + m) * sz);
            swap_elems(buf.data(), m, sz);
            if (fwrite(buf.data(), 1, buf.size(), f) != buf.size()) throw Error("netCDF writer: short write to " + fname);
        }
        static const uint8_t zeros[4] = {0, 0, 0, 0};
        const uint64_t padBytes = vsize[i] - n * sz;
        if (padBytes && fwrite(zeros, 1, padBytes, f) != padBytes) throw Error("netCDF writer: short write to " + fname);
    }
}

}  // namespace nc
