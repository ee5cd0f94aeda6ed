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

// ---- the reference's own coordinate sets (src/sample/coordinate_set.cpp): a frame reduced to a selection, translated, then
// changed to the spherical / cylindrical representation the multipole devices stage (:278-315) ----
#include "sample/atomselection.hpp"
#include "sample/coordinate_set.hpp"
extern "C" {
// xyz: double [natoms][3] (a Frame); sel: nsel atom indices; trans: translation applied to the set; repr 10 cartesian,
// 20 spherical (r, phi, theta), 30 cylindrical (r, phi, z) on `axis`.  out: double [nsel][3]
void ref_coordinate_set(const double *xyz, size_t natoms, const size_t *sel, size_t nsel, const double trans[3], int repr,
                        const double axis[3], double *out) {
    Frame fr;
    fr.number_of_atoms = natoms;
    for (size_t i = 0; i < natoms; i++) {
        fr.x.push_back(xyz[3 * i]);
        fr.y.push_back(xyz[3 * i + 1]);
        fr.z.push_back(xyz[3 * i + 2]);
    }
    IndexAtomselection s(std::vector<size_t>(sel, sel + nsel));
    CartesianCoordinateSet cs(fr, &s);
    cs.translate(CartesianCoor3D(trans[0], trans[1], trans[2]));
    CoordinateSet *res = &cs;
    SphericalCoordinateSet *sph = NULL;
    CylindricalCoordinateSet *cyl = NULL;
    if (repr == 20) res = sph = new SphericalCoordinateSet(cs);
    if (repr == 30) res = cyl = new CylindricalCoordinateSet(cs, CartesianCoor3D(axis[0], axis[1], axis[2]));
    for (size_t i = 0; i < res->size(); i++) {
        out[3 * i] = res->c1[i];
        out[3 * i + 1] = res->c2[i];
        out[3 * i + 2] = res->c3[i];
    }
    delete sph;
    delete cyl;
}
}
