// dcd.cpp — see dcd.hpp
#include "dcd.hpp"

#include <cstring>

#include "sassena_host.hpp"

namespace sassena {

namespace {
void rd(FILE *f, void *p, size_t n, const std::string &fn) {
    if (fread(p, 1, n, f) != n) throw Error("unexpected end of DCD file '" + fn + "'");
}
void seek(FILE *f, int64_t off, int whence, const std::string &fn) {
    if (fseeko(f, (off_t)off, whence) != 0) throw Error("seek failed in DCD file '" + fn + "'");
}
}  // namespace

bool DCDFrameset::detect(const std::string &fn) {
    FILE *f = fopen(fn.c_str(), "rb");
    if (!f) return false;
    unsigned char buf[92];
    size_t n = fread(buf, 1, 92, f);
    fclose(f);
    if (n != 92) return false;
    const unsigned char fp1[4] = {0x54, 0x00, 0x00, 0x00};
    const unsigned char fp2[4] = {0x43, 0x4f, 0x52, 0x44};  // "CORD"
    return memcmp(&buf[0], fp1, 4) == 0 && memcmp(&buf[88], fp1, 4) == 0 && memcmp(&buf[4], fp2, 4) == 0;
}

DCDFrameset::DCDFrameset(const std::string &fn) : filename_(fn) {
    if (!detect(fn)) throw Error("file '" + fn + "' appears to not be a DCD file. INT Magic failed");
    f_ = fopen(fn.c_str(), "rb");
    if (!f_) throw Error("cannot open '" + fn + "'");
    try {
        parse_header(fn);
    } catch (...) {  // the destructor does not run for a constructor that throws
        fclose(f_);
        f_ = nullptr;
        throw;
    }
}

void DCDFrameset::parse_header(const std::string &fn) {
    // DCDHeader (frames.hpp:152-164): headsize, fingerprint, number_of_frames, dummy1, timesteps_between_frames,
    // buf1[24], size_of_timestep, flag_ext_block1, flag_ext_block2
    unsigned char hdr[56];
    rd(f_, hdr, sizeof(hdr), fn);
    int32_t nof;
    memcpy(&nof, hdr + 8, 4);
    memcpy(&flag_ext_block1, hdr + 48, 4);
    memcpy(&flag_ext_block2, hdr + 52, 4);
    seek(f_, 23 * (int64_t)sizeof(int32_t), SEEK_SET, fn);
    int32_t marker;
    rd(f_, &marker, 4, fn);  // title block
    seek(f_, marker, SEEK_CUR, fn);
    rd(f_, &marker, 4, fn);
    uint32_t noa;
    rd(f_, &marker, 4, fn);
    rd(f_, &noa, 4, fn);
    rd(f_, &marker, 4, fn);
    number_of_frames = (size_t)nof;
    number_of_atoms = (size_t)noa;
    init_byte_pos = ftello(f_);
    auto rel = [&]() { return (int64_t)ftello(f_) - init_byte_pos; };
    if (flag_ext_block1) {
        rd(f_, &marker, 4, fn);
        block1_byte_offset = rel();
        seek(f_, marker, SEEK_CUR, fn);
        rd(f_, &marker, 4, fn);
    } else {
        block1_byte_offset = rel();
    }
    int64_t *offs[3] = {&x_byte_offset, &y_byte_offset, &z_byte_offset};
    for (int c = 0; c < 3; c++) {
        rd(f_, &marker, 4, fn);
        *offs[c] = rel();
        seek(f_, (int64_t)number_of_atoms * sizeof(float), SEEK_CUR, fn);
        rd(f_, &marker, 4, fn);
    }
    if (flag_ext_block2) {
        rd(f_, &marker, 4, fn);
        block2_byte_offset = rel();
        seek(f_, marker, SEEK_CUR, fn);
        rd(f_, &marker, 4, fn);
    } else {
        block2_byte_offset = rel();
    }
    block_size_byte = rel();
    // The reference trusts the header's frame count (frames.cpp:261-268) and reads stale data for frames the file does not
    // hold; a count that is negative or larger than the file can hold is refused here, before the index is sized by it.
    seek(f_, 0, SEEK_END, fn);
    const int64_t file_size = (int64_t)ftello(f_);
    if (nof < 0 || block_size_byte <= 0 || (int64_t)nof > (file_size - init_byte_pos) / block_size_byte)
        throw Error("DCD header of '" + fn + "' announces " + std::to_string(nof) + " frames, the file holds " +
                    std::to_string(block_size_byte > 0 ? (file_size - init_byte_pos) / block_size_byte : 0));
    // generate_index (frames.cpp:261-268)
    for (size_t i = 0; i < number_of_frames; ++i) frameset_index_.push_back((int64_t)i * block_size_byte + init_byte_pos);
    buf_.resize(number_of_atoms);
}

DCDFrameset::~DCDFrameset() {
    if (f_) fclose(f_);
}

void DCDFrameset::trim_index(size_t first, size_t last, bool last_set, size_t stride) {
    if (stride == 0) throw Error("frameset stride must be >= 1");
    std::vector<int64_t> lfo;
    for (size_t i = 0; i < frameset_index_.size(); ++i) {
        if (i < first) continue;
        if (last_set && (i > last)) break;
        if ((i % stride) == 0) lfo.push_back(frameset_index_[i]);
    }
    frameset_index_ = lfo;
    number_of_frames = frameset_index_.size();
}

void DCDFrameset::read_frame(size_t framenumber, float *xyz, double *unitcell) {
    if (framenumber >= frameset_index_.size()) throw Error("DCD frame number out of range");
    const int64_t base = frameset_index_[framenumber];
    if (unitcell) {
        if (flag_ext_block1) {
            seek(f_, base + block1_byte_offset, SEEK_SET, filename_);
            rd(f_, unitcell, 6 * sizeof(double), filename_);
        } else {
            for (int i = 0; i < 6; i++) unitcell[i] = 0.0;
        }
    }
    const int64_t offs[3] = {x_byte_offset, y_byte_offset, z_byte_offset};
    for (int c = 0; c < 3; c++) {
        seek(f_, base + offs[c], SEEK_SET, filename_);
        rd(f_, buf_.data(), number_of_atoms * sizeof(float), filename_);
        for (size_t i = 0; i < number_of_atoms; i++) xyz[3 * i + c] = buf_[i];
    }
}

void DCDFrameset::read_frames(size_t first, size_t count, float *out) {
    for (size_t i = 0; i < count; i++) read_frame(first + i, out + i * number_of_atoms * 3);
}

// ---- writer (coordinate_writer.cpp:37-144) ----
void DCDCoordinateWriter::init() {
    FILE *out = fopen(file_.c_str(), "wb");
    if (!out) throw Error("cannot create '" + file_ + "'");
    auto w32 = [&](int32_t v) { fwrite(&v, 4, 1, out); };
    w32(4 * 21);  // head size
    fwrite("CORD", 1, 4, out);
    w32((int32_t)blocks_);  // number of frames
    w32(0);
    w32(1);                              // timesteps between frames
    for (int i = 0; i < 6; i++) w32(0);  // 6..11
    float dt = 1;
    fwrite(&dt, 4, 1, out);              // size of timestep
    w32(1);                              // optional block 1 (unit cell) on
    w32(0);                              // optional block 2 off
    for (int i = 0; i < 7; i++) w32(0);
    w32(24);      // != 0: CHARMM format
    w32(4 * 21);  // closing header marker
    w32(4);       // title block: one int, value 0
    w32(0);
    w32(4);
    w32(4);
    w32((int32_t)entries_);  // number of atoms
    w32(4);
    fclose(out);
}

void DCDCoordinateWriter::prepare() {
    FILE *in = fopen(file_.c_str(), "rb");
    if (!in) throw Error("cannot open '" + file_ + "'");
    fseeko(in, 0, SEEK_END);
    data_offset_ = ftello(in);
    fclose(in);
}

void DCDCoordinateWriter::write(const float *data, size_t blockoffset, size_t myblocks) {
    FILE *out = fopen(file_.c_str(), "r+b");
    if (!out) throw Error("cannot open '" + file_ + "' for writing");
    const size_t blockbytesize = 2 * sizeof(int32_t) + 6 * sizeof(double) + 3 * 2 * sizeof(int32_t) + 3 * entries_ * sizeof(float);
    fseeko(out, (off_t)(data_offset_ + (int64_t)blockoffset * (int64_t)blockbytesize), SEEK_SET);
    std::vector<float> buf(entries_);
    for (size_t i = 0; i < myblocks; ++i) {
        int32_t marker = 6 * sizeof(double);
        double emptycell[6] = {0, 0, 0, 0, 0, 0};
        fwrite(&marker, 4, 1, out);
        fwrite(emptycell, sizeof(double), 6, out);
        fwrite(&marker, 4, 1, out);
        for (size_t k = 0; k < 3; ++k) {
            for (size_t j = 0; j < entries_; ++j) buf[j] = data[i * entries_ * 3 + j * 3 + k];
            marker = (int32_t)(entries_ * sizeof(float));
            fwrite(&marker, 4, 1, out);
            fwrite(buf.data(), sizeof(float), entries_, out);
            fwrite(&marker, 4, 1, out);
        }
    }
    fclose(out);
}

}  // namespace sassena
