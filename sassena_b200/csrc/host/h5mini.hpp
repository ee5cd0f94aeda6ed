// h5mini.hpp — a minimal HDF5 container writer/reader for the signal file ("next" row).
// The reference writes signal.h5 through libhdf5 (src/services/file_writer_service.cpp:44-171 layout, :314-484 append,
// :504-515 client side): datasets qvectors [N,3], fqt [N,NF,2], fq0/fq/fq2 [N,2], all float64, extendible along N and
// chunked by limits.signal.chunksize, plus group meta/{rawconfig (char array), config, database (C strings)}.
// libhdf5 is not available in this image, so the container is written from the HDF5 file-format specification in the
// subset a default libhdf5-1.8 H5Fcreate/H5Dcreate produces and any libhdf5 reads:
//   superblock version 0, version-1 object headers, symbol-table groups (v1 B-tree + local heap + SNOD),
//   dataspace v1, datatype v1 (IEEE f64 LE / 1-byte signed integer / fixed C string), fill-value v2,
//   data layout v3 contiguous or chunked with a v1 "raw data chunk" B-tree (no filters).
// The reader handles the same subset (+ layout v1/v2, userblocks, header continuation blocks) and is pinned in the tests
// against a file written by the real library (a MATLAB v7.3 file shipped with scipy); it then checks the writer.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace sassena {
namespace h5 {

constexpr uint64_t UNLIMITED = ~(uint64_t)0;

struct Dataset {
    enum Kind { F64 = 0, CHAR = 1, STRING = 2 };
    std::string name;
    Kind kind = F64;
    std::vector<uint64_t> dims;     // current extent (empty = scalar)
    std::vector<uint64_t> maxdims;  // empty = same as dims (not extendible); UNLIMITED allowed (needs chunks)
    std::vector<uint64_t> chunk;    // empty = contiguous layout
    std::vector<double> f64;        // row-major values (F64)
    std::string bytes;              // CHAR: dims[0] bytes; STRING: the text (stored null-terminated, scalar dataspace)
};

struct Group {
    std::string name;
    std::vector<Dataset> datasets;
    std::vector<Group> groups;
    const Dataset *find(const std::string &n) const;
    const Group *find_group(const std::string &n) const;
};

// serialise `root` (its name is ignored) into a complete HDF5 file image / file
std::vector<uint8_t> serialize(const Group &root);
void write_file(const std::string &path, const Group &root);  // writes path + ".tmp", then renames
// parse a file (throws sassena::Error on anything outside the supported subset)
Group read_file(const std::string &path);
Group parse(const std::vector<uint8_t> &image);

}  // namespace h5

// HDF5WriterService / HDF5WriterClient stand-in (file_writer_service.cpp): collects the rows the devices hand over and
// keeps signal.h5 up to date.  init(): a fresh file, or -- when the file exists and its fqt extent matches NF -- the
// reference's resume (sassena.cpp:270-305): rows already present are kept and their q-vectors reported as done.
class SignalFileH5 {
    std::string path_;
    size_t NF_ = 0, chunksize_ = 10000;
    bool fqt_ = true, fq0_ = true, fq_ = true, fq2_ = true;
    std::string rawconfig_, config_, database_;
    std::vector<double> q_, vfqt_, vfq0_, vfq_, vfq2_;

   public:
    SignalFileH5(const std::string &path, size_t NF, size_t chunksize, bool fqt, bool fq0, bool fq, bool fq2);
    void set_meta(const std::string &rawconfig, const std::string &config, const std::string &database);
    // returns the q-vectors already stored (3 doubles each); empty for a new file
    std::vector<double> init();
    void write(const double q[3], const double *fqt, const double fq[2], const double fq2[2]);
    void flush() const;
    size_t rows() const { return q_.size() / 3; }
};

}  // namespace sassena
