// NetCDF.h -- minimal reader / writer of the netCDF *classic* file family (CDF-1, CDF-2 "64-bit offset", CDF-5 "64-bit
// data"), enough for DSSTNE's dataset and network files ("next" row 1 of SURVEY 8f).  Host only, no libnetcdf.
//
// The reference reads and writes through netcdf-cxx4 (E/NNTypes.cpp:1081-1418, 2218-2385; E/NNNetwork.cpp:1936-1970;
// U/NetCDFhelper.cpp:332-416), whose default on-disk format is netCDF-4 = HDF5.  Such containers are recognised by their
// signature and read by HDF5.cpp (contiguous / compact fixed-size variables and attributes of the netCDF atomic types;
// chunked or compressed variables are refused by name with the `nccopy` line that rewrites them).  Files written here
// are CDF-5 (or CDF-2 on request, for tools that only read classic types) and open with any netCDF >= 4.4; CDF-5 keeps
// the unsigned and 64-bit variable types DSSTNE uses.
//
// Layout handled: fixed-size variables only (DSSTNE never uses the record dimension), any of the 11 atomic types,
// global and per-variable attributes.
#pragma once

#include <cstdint>
#include <cstdio>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace nc {

enum Type { NC_BYTE = 1, NC_CHAR = 2, NC_SHORT = 3, NC_INT = 4, NC_FLOAT = 5, NC_DOUBLE = 6, NC_UBYTE = 7, NC_USHORT = 8, NC_UINT = 9,
            NC_INT64 = 10, NC_UINT64 = 11 };

size_t type_size(Type t);
const char* type_name(Type t);

struct Error : public std::runtime_error {
    explicit Error(const std::string& m) : std::runtime_error(m) {}
};

struct Att {
    std::string name;
    Type type;
    uint64_t nelems;
    std::vector<uint8_t> data;                       // host byte order, nelems * type_size(type) bytes
    std::string as_string() const;                   // NC_CHAR attributes
    double as_double(size_t i = 0) const;            // any numeric attribute
    int64_t as_int(size_t i = 0) const;
};

struct Dim {
    std::string name;
    uint64_t size;
};

struct Var {
    std::string name;
    std::vector<uint32_t> dimids;
    std::vector<Att> atts;
    Type type;
    uint64_t vsize;                                   // bytes in the file (padded to 4)
    uint64_t begin;                                   // file offset of the data
    uint64_t nelems;                                  // product of the dimension sizes
    bool littleEndian = false;                        // netCDF-4 / HDF5 variables are stored in the writer's byte order (classic: big-endian)
    bool hasInline = false;                           // HDF5 compact layout (or a never-written variable): the bytes live in the header
    std::vector<uint8_t> inlineData;
    // HDF5 chunked layout (what netCDF-4 uses for compressed variables and variables over an unlimited dimension): HDF5.cpp reads it
    struct Chunk { uint64_t addr; uint32_t bytes; uint32_t filterMask; std::vector<uint64_t> start; };
    bool chunked = false;
    uint32_t chunkElemBytes = 0;
    std::vector<uint64_t> shape, chunkShape;          // in elements
    std::vector<uint32_t> filters;                    // the pipeline in write order: 1 deflate, 2 shuffle, 3 fletcher32
    std::vector<uint8_t> fill;                        // one element: what never-written chunks read as (empty: zero)
    std::vector<Chunk> chunks;
};

class File {
public:
    explicit File(const std::string& fname);          // parses the header; throws nc::Error
    int version() const { return _version; }          // 1, 2 or 5 (classic family), 4 (netCDF-4 / HDF5, read only: HDF5.cpp)
    const std::vector<Dim>& dims() const { return _dims; }
    const std::vector<Att>& atts() const { return _atts; }
    const std::vector<Var>& vars() const { return _vars; }
    const Att* att(const std::string& name) const;
    const Dim* dim(const std::string& name) const;
    const Var* var(const std::string& name) const;
    // whole variable, converted element-wise to T (integer <-> integer / float conversions as static_cast)
    template <typename T> void read(const Var& v, std::vector<T>& out) const;
    std::string describe() const;                     // ncdump -h like text

private:
    std::string _fname;
    int _version;
    uint64_t _numrecs;
    std::vector<Dim> _dims;
    std::vector<Att> _atts;
    std::vector<Var> _vars;
    void read_raw(const Var& v, std::vector<uint8_t>& bytes) const;   // file (big-endian) bytes of the variable
};

// Builds a file in memory order: define everything, then write().  Variable data is borrowed until write().
class Writer {
public:
    explicit Writer(int version = 5);                 // 5 = CDF-5, 2 = CDF-2 (classic types only)
    void add_dim(const std::string& name, uint64_t size);
    void put_att(const std::string& name, const std::string& value);
    void put_att(const std::string& name, Type type, double value);
    void put_att_u64(const std::string& name, Type type, uint64_t value);
    // one-dimensional variable over `dim`; `data` holds nelems elements of the host type matching `type`
    void add_var(const std::string& name, Type type, const std::string& dim, const void* data);
    void write(const std::string& fname) const;

private:
    struct V { std::string name; Type type; uint32_t dimid; const void* data; };
    int _version;
    std::vector<Dim> _dims;
    std::vector<Att> _atts;
    std::vector<V> _vars;
};

namespace hdf5 {
// netCDF-4 container -> the same dimension / attribute / variable tables the classic parser fills (HDF5.cpp)
void parse(FILE* f, const std::string& fname, std::vector<Dim>& dims, std::vector<Att>& atts, std::vector<Var>& vars);
// bytes of a chunked variable in row-major order, filters undone, in the file's byte order
void read_chunked(FILE* f, const std::string& fname, const Var& v, std::vector<uint8_t>& bytes);
}

}  // namespace nc
