// xdr_traj.hpp — GROMACS XTC / TRR trajectory readers feeding the stager ("next" row).
// The reference reads both through the vendored xdrfile library (src/sample/frames.cpp:592-858: read_xtc / read_trr,
// one index entry per frame start, coordinates scaled nm -> Angstrom by 10.0).  xdrfile is third-party code that is not
// part of this repository; the readers here are written from the published file formats (XDR big-endian primitives,
// the xtc "3dfcoord" integer compression, the trn header) and pinned against xdrfile itself in the tests: oracle/Makefile
// builds the reference's vendored copy into oracle/_ref/libxdrfile_ref.so, tests/golden/make_xdr_golden.py writes
// trajectories with it, and the readers must reproduce xdrfile's own decoding bit for bit.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace sassena {

class XdrFile {  // big-endian 4-byte-aligned primitives on a FILE*
    FILE *f_ = nullptr;
    std::string name_;
    int64_t size_ = 0;

   public:
    explicit XdrFile(const std::string &fn);
    ~XdrFile();
    XdrFile(const XdrFile &) = delete;
    XdrFile &operator=(const XdrFile &) = delete;
    int64_t tell() const;
    int64_t size() const { return size_; }
    void seek(int64_t pos);
    bool read_i32(int32_t &v);  // false at end of file
    bool read_f32(float &v);
    bool read_f64(double &v);
    bool read_bytes(void *dst, size_t n);  // n payload bytes followed by padding to a multiple of 4
    bool skip(int64_t n);
    const std::string &name() const { return name_; }
};

// common part: frame index, first/last/stride trimming (FileFrameset::trim_index, frames.cpp:224-245)
class XdrFrameset {
   protected:
    XdrFile file_;
    std::vector<int64_t> frameset_index_;
    explicit XdrFrameset(const std::string &fn) : file_(fn) {}

   public:
    size_t number_of_frames = 0, number_of_atoms = 0;
    virtual ~XdrFrameset() {}
    void trim_index(size_t first, size_t last, bool last_set, size_t stride);
    // xyz: [number_of_atoms][3] in Angstrom, each component (float)(10.0 * (double)nm) as the reference's Frame + stager
    // narrowing produce it; box (9 doubles, Angstrom) may be NULL
    virtual void read_frame(size_t framenumber, float *xyz, double *box = nullptr) = 0;
};

class XTCFrameset : public XdrFrameset {
    std::vector<int32_t> ints_;
    std::vector<uint8_t> packed_;
    // reads the frame at the current position into nm coordinates; returns false at a clean end of file
    bool read_frame_nm(float *xyz_nm, float *box_nm, bool decode);

   public:
    explicit XTCFrameset(const std::string &fn);
    static bool detect(const std::string &fn);
    void read_frame(size_t framenumber, float *xyz, double *box = nullptr) override;
};

class TRRFrameset : public XdrFrameset {
    struct Header {
        int32_t ir_size, e_size, box_size, vir_size, pres_size, top_size, sym_size, x_size, v_size, f_size, natoms, step, nre;
        bool is_double;
    };
    bool read_header(Header &h);
    bool read_frame_nm(float *xyz_nm, double *box_nm, bool decode);

   public:
    explicit TRRFrameset(const std::string &fn);
    static bool detect(const std::string &fn);
    void read_frame(size_t framenumber, float *xyz, double *box = nullptr) override;
};

}  // namespace sassena
