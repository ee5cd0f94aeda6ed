// xdr_traj.cpp — GROMACS XTC / TRR readers written from the file formats (see xdr_traj.hpp for the reference call sites).
//
// File layouts (all items are XDR: big-endian, 4-byte aligned):
//   XTC frame : i32 magic(1995) | i32 natoms | i32 step | f32 time | f32 box[9] | i32 natoms |
//               natoms <= 9 : f32 x[3*natoms]
//               otherwise   : f32 precision | i32 lo[3] | i32 hi[3] | i32 smallidx | i32 nbytes | bytes[nbytes] (+pad)
//   TRR frame : i32 magic(1993) | i32 slen(13) | string "GMX_trn_file" | i32 ir,e,box,vir,pres,top,sym,x,v,f sizes |
//               i32 natoms | i32 step | i32 nre | real t | real lambda | box | vir | pres | x | v | f      (real = f32/f64)
// The packed XTC payload is a bit stream (most significant bit first).  Every atom is either a "large" triple — three
// integers in [0, hi-lo] packed as one mixed-radix number (or three plain bit fields when a range exceeds 24 bits) — or,
// inside a run announced by a 1-bit flag + 5-bit run code, a "small" triple: offsets from the previous atom packed as one
// mixed-radix number with radix M[smallidx], M[k] ~ 2^(k/3).  The first small atom of a run is swapped with the large
// atom before it (water: O stored after the first H).  Coordinates are integer * (1/precision) evaluated in float.
#include "xdr_traj.hpp"

#include <cmath>
#include <cstring>

#include "sassena_host.hpp"

namespace sassena {

// ---------------------------------------------------------------- XDR primitives
XdrFile::XdrFile(const std::string &fn) : name_(fn) {
    f_ = fopen(fn.c_str(), "rb");
    if (!f_) throw Error("Unable to open file: " + fn);
    fseeko(f_, 0, SEEK_END);
    size_ = ftello(f_);
    fseeko(f_, 0, SEEK_SET);
}
XdrFile::~XdrFile() {
    if (f_) fclose(f_);
}
int64_t XdrFile::tell() const { return ftello(f_); }
void XdrFile::seek(int64_t pos) { fseeko(f_, pos, SEEK_SET); }

static inline uint32_t be32(const unsigned char *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

bool XdrFile::read_i32(int32_t &v) {
    unsigned char b[4];
    if (fread(b, 1, 4, f_) != 4) return false;
    v = (int32_t)be32(b);
    return true;
}
bool XdrFile::read_f32(float &v) {
    int32_t i;
    if (!read_i32(i)) return false;
    memcpy(&v, &i, 4);
    return true;
}
bool XdrFile::read_f64(double &v) {
    unsigned char b[8];
    if (fread(b, 1, 8, f_) != 8) return false;
    uint64_t u = ((uint64_t)be32(b) << 32) | be32(b + 4);
    memcpy(&v, &u, 8);
    return true;
}
bool XdrFile::read_bytes(void *dst, size_t n) {
    if (n && fread(dst, 1, n, f_) != n) return false;
    const size_t pad = (4 - n % 4) % 4;
    unsigned char scratch[4];
    if (pad && fread(scratch, 1, pad, f_) != pad) return false;
    return true;
}
bool XdrFile::skip(int64_t n) {
    const int64_t target = tell() + n;
    if (target > size_) return false;
    seek(target);
    return true;
}

void XdrFrameset::trim_index(size_t first, size_t last, bool last_set, size_t stride) {  // frames.cpp:224-245
    std::vector<int64_t> kept;
    for (size_t i = 0; i < frameset_index_.size(); i++) {
        if (i < first) continue;
        if (last_set && i > last) continue;
        if (stride > 1 && i % stride != 0) continue;
        kept.push_back(frameset_index_[i]);
    }
    frameset_index_.swap(kept);
    number_of_frames = frameset_index_.size();
}

// ---------------------------------------------------------------- XTC
namespace {

constexpr int kXtcMagic = 1995;
constexpr int kFirstIdx = 9;   // smallest radix index a file may use
constexpr int kNumRadix = 73;  // radix table entries 0..72

// Radix table of the format: M[k] = floor(2^(k/3)) for k >= 9, with the three values the original encoder's table
// carries (k = 37, 57, 69 differ from the formula and are part of the format).
struct RadixTable {
    int m[kNumRadix];
    RadixTable() {
        for (int k = 0; k < kNumRadix; k++) {
            if (k < kFirstIdx) {
                m[k] = 0;
                continue;
            }
            // largest r with r^3 <= 2^k
            const unsigned __int128 target = (unsigned __int128)1 << k;
            uint64_t r = (uint64_t)std::floor(std::exp2(k / 3.0));
            while ((unsigned __int128)r * r * r > target) r--;
            while ((unsigned __int128)(r + 1) * (r + 1) * (r + 1) <= target) r++;
            m[k] = (int)r;
        }
        m[37] = 5060;
        m[57] = 524287;
        m[69] = 8388607;
    }
};
const RadixTable kRadix;

int bits_for(uint32_t range) {  // bits needed to store values 0..range (range >= 1)
    int n = 0;
    uint64_t cap = 1;
    while (range >= cap && n < 32) {
        n++;
        cap <<= 1;
    }
    return n;
}

int bits_for_product(const uint32_t r[3]) {  // bits of the mixed-radix number with digits < r[0], r[1], r[2]
    unsigned __int128 p = (unsigned __int128)r[0] * r[1] * r[2];
    // the encoder counts whole bytes below the top byte plus the bits needed for the value of the top byte
    int nbytes = 0;
    unsigned __int128 t = p;
    while (t >> 8) {
        t >>= 8;
        nbytes++;
    }
    int n = 0;
    unsigned top = (unsigned)t, cap = 1;
    while (top >= cap) {
        n++;
        cap *= 2;
    }
    return n + 8 * nbytes;
}

class BitStream {
    const uint8_t *p_;
    size_t n_, pos_ = 0;
    uint64_t acc_ = 0;
    int have_ = 0;

   public:
    BitStream(const uint8_t *p, size_t n) : p_(p), n_(n) {}
    uint32_t take(int nbits) {  // nbits <= 32, most significant bit first; reads zeros past the end
        while (have_ < nbits) {
            acc_ = (acc_ << 8) | (pos_ < n_ ? p_[pos_] : 0);
            pos_++;
            have_ += 8;
        }
        have_ -= nbits;
        const uint64_t v = (acc_ >> have_) & ((nbits >= 64) ? ~0ull : ((1ull << nbits) - 1));
        acc_ &= (have_ ? ((1ull << have_) - 1) : 0);
        return (uint32_t)v;
    }
    // mixed-radix triple stored in nbits: the stream holds the number's bytes least significant first, the last
    // (most significant) chunk holding the leftover nbits % 8 bits
    void take_triple(int nbits, const uint32_t radix[3], int out[3]) {
        unsigned __int128 v = 0;
        int shift = 0;
        while (nbits > 8) {
            v |= (unsigned __int128)take(8) << shift;
            shift += 8;
            nbits -= 8;
        }
        if (nbits > 0) v |= (unsigned __int128)take(nbits) << shift;
        out[2] = (int)(uint32_t)(v % radix[2]);
        v /= radix[2];
        out[1] = (int)(uint32_t)(v % radix[1]);
        v /= radix[1];
        out[0] = (int)(uint32_t)v;
    }
    bool overrun() const { return pos_ > n_ + 8; }
};

}  // namespace

XTCFrameset::XTCFrameset(const std::string &fn) : XdrFrameset(fn) {
    int32_t magic, natoms;
    if (!file_.read_i32(magic) || magic != kXtcMagic || !file_.read_i32(natoms) || natoms < 0)
        throw Error("file '" + fn + "' appears not to be a XTC file");
    number_of_atoms = (size_t)natoms;
    file_.seek(0);
    // generate_index (frames.cpp:615-658): one entry per frame that reads back completely
    while (true) {
        const int64_t pos = file_.tell();
        if (!read_frame_nm(nullptr, nullptr, false)) break;
        frameset_index_.push_back(pos);
    }
    number_of_frames = frameset_index_.size();
}

bool XTCFrameset::detect(const std::string &fn) {
    FILE *f = fopen(fn.c_str(), "rb");
    if (!f) return false;
    unsigned char b[16];
    const bool ok = fread(b, 1, 16, f) == 16 && be32(b) == (uint32_t)kXtcMagic;
    fclose(f);
    return ok;
}

bool XTCFrameset::read_frame_nm(float *xyz, float *box, bool decode) {
    int32_t magic, natoms, step, lsize;
    float time;
    if (!file_.read_i32(magic)) return false;  // clean end of file
    if (magic != kXtcMagic) {
        if (!decode) return false;  // the reference stops indexing at the first frame that does not read
        throw Error("XTC magic number mismatch in " + file_.name());
    }
    if (!file_.read_i32(natoms) || !file_.read_i32(step) || !file_.read_f32(time)) return false;
    for (int i = 0; i < 9; i++) {
        float b;
        if (!file_.read_f32(b)) return false;
        if (box) box[i] = b;
    }
    if (!file_.read_i32(lsize)) return false;
    if (lsize < 0 || (size_t)lsize != number_of_atoms || (size_t)natoms != number_of_atoms) {
        if (!decode) return false;
        throw Error("XTC frame with " + std::to_string(lsize) + " atoms in a trajectory of " + std::to_string(number_of_atoms));
    }
    if (lsize <= 9) {  // stored as plain floats
        for (int i = 0; i < 3 * lsize; i++) {
            float v;
            if (!file_.read_f32(v)) return false;
            if (decode) xyz[i] = v;
        }
        return true;
    }
    float precision;
    int32_t lo[3], hi[3], smallidx, nbytes;
    if (!file_.read_f32(precision)) return false;
    for (int c = 0; c < 3; c++)
        if (!file_.read_i32(lo[c])) return false;
    for (int c = 0; c < 3; c++)
        if (!file_.read_i32(hi[c])) return false;
    if (!file_.read_i32(smallidx) || !file_.read_i32(nbytes) || nbytes < 0) return false;
    if (!decode) return file_.skip(((int64_t)nbytes + 3) / 4 * 4);
    if (smallidx < kFirstIdx || smallidx >= kNumRadix) throw Error("corrupt XTC frame (radix index) in " + file_.name());
    packed_.resize((size_t)nbytes);
    if (!file_.read_bytes(packed_.data(), (size_t)nbytes)) return false;

    uint32_t range[3];
    for (int c = 0; c < 3; c++) range[c] = (uint32_t)hi[c] - (uint32_t)lo[c] + 1u;  // modulo 2^32: no UB on a corrupt header
    int wide_bits[3] = {0, 0, 0}, large_bits = 0;
    if ((range[0] | range[1] | range[2]) > 0xffffffu) {
        for (int c = 0; c < 3; c++) wide_bits[c] = bits_for(range[c]);
    } else {
        large_bits = bits_for_product(range);
    }
    const float inv_precision = (float)(1.0 / precision);
    BitStream bs(packed_.data(), packed_.size());
    auto emit = [&](float *&dst, const int v[3]) {
        for (int c = 0; c < 3; c++) *dst++ = v[c] * inv_precision;
    };
    float *dst = xyz;
    int run = 0;
    size_t atom = 0;
    const size_t n = (size_t)lsize;
    while (atom < n) {
        int cur[3];
        if (large_bits == 0) {
            for (int c = 0; c < 3; c++) cur[c] = (int)bs.take(wide_bits[c]);
        } else {
            bs.take_triple(large_bits, range, cur);
        }
        for (int c = 0; c < 3; c++) cur[c] = (int)((uint32_t)cur[c] + (uint32_t)lo[c]);
        atom++;
        int step_idx = 0;  // change of the small radix after this group: -1, 0, +1
        if (bs.take(1)) {
            const int code = (int)bs.take(5);
            step_idx = code % 3 - 1;
            run = code - code % 3;  // number of small coordinates (3 per atom) that follow; sticky across groups
        }
        if (run > 0) {
            if (atom + (size_t)run / 3 > n) throw Error("corrupt XTC frame (run past the last atom) in " + file_.name());
            const int radix = kRadix.m[smallidx];
            const uint32_t small_r[3] = {(uint32_t)radix, (uint32_t)radix, (uint32_t)radix};
            const int half = radix / 2;
            int prev[3] = {cur[0], cur[1], cur[2]};
            for (int k = 0; k < run; k += 3) {
                int s[3];
                bs.take_triple(smallidx, small_r, s);
                for (int c = 0; c < 3; c++) s[c] = (int)((uint32_t)s[c] + (uint32_t)prev[c] - (uint32_t)half);
                atom++;
                emit(dst, s);
                if (k == 0) emit(dst, cur);  // the first small atom was stored after the large one: it comes out first
                for (int c = 0; c < 3; c++) prev[c] = s[c];  // offsets chain from small atom to small atom
            }
        } else {
            emit(dst, cur);
        }
        smallidx += step_idx;
        if (smallidx < kFirstIdx || smallidx >= kNumRadix) throw Error("corrupt XTC frame (radix index) in " + file_.name());
        if (bs.overrun()) throw Error("corrupt XTC frame (payload too short) in " + file_.name());
    }
    return true;
}

void XTCFrameset::read_frame(size_t framenumber, float *xyz, double *box) {
    if (framenumber >= frameset_index_.size()) throw Error("XTC frame number out of range");
    file_.seek(frameset_index_[framenumber]);
    float b[9];
    if (!read_frame_nm(xyz, b, true)) throw Error("short read in " + file_.name());
    // nm -> Angstrom in double, then the stager's narrowing to float (frames.cpp:705-714, data_stager.cpp:111-113)
    for (size_t i = 0; i < 3 * number_of_atoms; i++) xyz[i] = (float)(10.0 * (double)xyz[i]);
    if (box)
        for (int i = 0; i < 9; i++) box[i] = 10.0 * (double)b[i];
}

// ---------------------------------------------------------------- TRR
namespace {
constexpr int kTrrMagic = 1993;
const char kTrrVersion[] = "GMX_trn_file";
}  // namespace

bool TRRFrameset::read_header(Header &h) {
    int32_t magic, slen, n;
    if (!file_.read_i32(magic)) return false;
    if (magic != kTrrMagic) return false;
    if (!file_.read_i32(slen) || slen != (int32_t)sizeof(kTrrVersion)) return false;
    if (!file_.read_i32(n) || n < 0 || n > 255) return false;
    char tag[256];
    if (!file_.read_bytes(tag, (size_t)n)) return false;
    int32_t *fields[] = {&h.ir_size, &h.e_size, &h.box_size, &h.vir_size, &h.pres_size, &h.top_size, &h.sym_size,
                         &h.x_size,  &h.v_size, &h.f_size,   &h.natoms};
    for (int32_t *p : fields)
        if (!file_.read_i32(*p)) return false;
    // width of a real: from the first block that is present
    int64_t width = 0;
    const int64_t per_atom = 3 * (int64_t)h.natoms;  // 64-bit: a corrupt atom count must not overflow
    if (h.box_size)
        width = h.box_size / 9;
    else if (h.natoms > 0 && h.x_size)
        width = h.x_size / per_atom;
    else if (h.natoms > 0 && h.v_size)
        width = h.v_size / per_atom;
    else if (h.natoms > 0 && h.f_size)
        width = h.f_size / per_atom;
    if (width != 4 && width != 8) return false;
    h.is_double = width == 8;
    if (!file_.read_i32(h.step) || !file_.read_i32(h.nre)) return false;
    return file_.skip(2 * (int64_t)width);  // time, lambda
}

bool TRRFrameset::read_frame_nm(float *xyz, double *box, bool decode) {
    Header h;
    if (!read_header(h)) return false;
    if ((size_t)h.natoms != number_of_atoms) {
        if (!decode) return false;
        throw Error("TRR frame with " + std::to_string(h.natoms) + " atoms in a trajectory of " + std::to_string(number_of_atoms));
    }
    const int width = h.is_double ? 8 : 4;
    auto read_real = [&](double &v) {
        if (h.is_double) return file_.read_f64(v);
        float f;
        if (!file_.read_f32(f)) return false;
        v = f;
        return true;
    };
    if (h.box_size) {
        for (int i = 0; i < 9; i++) {
            double v;
            if (!read_real(v)) return false;
            if (box) box[i] = (double)(float)v;  // xdrfile hands the box back as float
        }
    }
    if (h.vir_size && !file_.skip(9 * width)) return false;
    if (h.pres_size && !file_.skip(9 * width)) return false;
    const int64_t block = (int64_t)number_of_atoms * 3 * width;
    if (h.x_size) {
        if (decode) {
            for (size_t i = 0; i < 3 * number_of_atoms; i++) {
                double v;
                if (!read_real(v)) return false;
                xyz[i] = (float)v;
            }
        } else if (!file_.skip(block)) {
            return false;
        }
    } else if (decode) {
        for (size_t i = 0; i < 3 * number_of_atoms; i++) xyz[i] = 0.f;  // no positions in this frame
    }
    if (h.v_size && !file_.skip(block)) return false;
    if (h.f_size && !file_.skip(block)) return false;
    return true;
}

TRRFrameset::TRRFrameset(const std::string &fn) : XdrFrameset(fn) {
    Header h;
    if (!read_header(h) || h.natoms < 0) throw Error("file '" + fn + "' appears not to be a TRR file");
    number_of_atoms = (size_t)h.natoms;
    file_.seek(0);
    while (true) {  // generate_index (frames.cpp:763-806)
        const int64_t pos = file_.tell();
        if (!read_frame_nm(nullptr, nullptr, false)) break;
        frameset_index_.push_back(pos);
    }
    number_of_frames = frameset_index_.size();
}

bool TRRFrameset::detect(const std::string &fn) {
    try {
        XdrFile f(fn);
        int32_t magic, slen;
        return f.read_i32(magic) && magic == kTrrMagic && f.read_i32(slen) && slen == (int32_t)sizeof(kTrrVersion);
    } catch (const Error &) {
        return false;
    }
}

void TRRFrameset::read_frame(size_t framenumber, float *xyz, double *box) {
    if (framenumber >= frameset_index_.size()) throw Error("TRR frame number out of range");
    file_.seek(frameset_index_[framenumber]);
    double b[9] = {0};
    if (!read_frame_nm(xyz, b, true)) throw Error("short read in " + file_.name());
    for (size_t i = 0; i < 3 * number_of_atoms; i++) xyz[i] = (float)(10.0 * (double)xyz[i]);  // frames.cpp:840-846
    if (box)
        for (int i = 0; i < 9; i++) box[i] = 10.0 * b[i];
}

}  // namespace sassena
