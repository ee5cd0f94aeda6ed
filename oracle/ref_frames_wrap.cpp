// ref_frames_wrap.cpp — the REFERENCE's own trajectory readers (src/sample/frames.cpp: DCDFrameset, PDBFrameset, XTCFrameset,
// TRRFrameset with generate_index / trim_index / read_frame, over its vendored xdrfile) compiled where they lie into
// oracle/_ref/libparams_ref.so.  The .tnx index files (Boost text archives) cannot be read or written here; every frameset is
// indexed with the reference's own generate_index().  Test infrastructure: pins the product's readers (csrc/host/dcd.cpp,
// xdr_traj.cpp, the PDB framesets of control.cpp).
#include <string>
#include <vector>

#include "sample/frame.hpp"
#include "sample/frames.hpp"

namespace {
FileFrameset *open_frameset(const std::string &format, const std::string &file) {
    if (format == "dcd") return new DCDFrameset(file, 0);
    if (format == "pdb") return new PDBFrameset(file, 0);
    // (the (filename, offset) constructors of the XTC / TRR framesets leave p_xdrfile uninitialised, which generate_index then
    // tests against NULL; the default constructor + init() the header offers sets it)
    if (format == "xtc") {
        XTCFrameset *x = new XTCFrameset();
        x->init(file, 0);
        return x;
    }
    if (format == "trr") {
        TRRFrameset *x = new TRRFrameset();
        x->init(file, 0);
        return x;
    }
    return NULL;
}
}  // namespace

extern "C" {
// reads the frames kept by (first, last, last_set, stride) into out [nframes][natoms][3] (double); out == NULL: only counts.
// returns the number of frames kept; *natoms_out receives the number of atoms per frame
size_t ref_frames_read(const char *format, const char *file, size_t first, size_t last, int last_set, size_t stride,
                       size_t *natoms_out, double *out, size_t cap_frames) {
    FileFrameset *fs = open_frameset(format, file);
    if (!fs) return (size_t)-1;
    fs->generate_index();
    fs->trim_index(first, last, last_set != 0, stride);
    size_t n = fs->number_of_frames;
    // atoms per frame from the first frame read (PDBFrameset never sets number_of_atoms, frames.cpp:442-489)
    size_t na = 0;
    for (size_t f = 0; f < n; f++) {
        if (f > 0 && (!out || f >= cap_frames)) break;
        Frame fr;
        fs->read_frame(f, fr);
        if (f == 0) na = fr.x.size();
        if (!out || fr.x.size() != na) continue;
        for (size_t a = 0; a < na; a++) {
            out[(f * na + a) * 3] = fr.x[a];
            out[(f * na + a) * 3 + 1] = fr.y[a];
            out[(f * na + a) * 3 + 2] = fr.z[a];
        }
    }
    *natoms_out = na;
    delete fs;
    return n;
}
}
