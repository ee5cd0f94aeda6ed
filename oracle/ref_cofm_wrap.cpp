// ref_cofm_wrap.cpp — the REFERENCE's own rotational fit (Fit, src/sample/center_of_mass.cpp:53-157: mass-weighted correlation
// kernel, LAPACK dgesvd through the Boost.Bindings call, determinant correction, write-back through the selections) and its
// CenterOfMass, compiled where they lie into oracle/_ref/libparams_ref.so over the uBLAS / bindings shims (oracle/shim/boost/
// numeric/ublas, oracle/vendor/boost/numeric/bindings: dgesvd is a real LAPACK, the OpenBLAS scipy bundles).  Test
// infrastructure: pins the product's fitrot / fitrottrans alignments (csrc/host/coordinate_sets.cpp, Horn's quaternion form).
#include <string>
#include <vector>

#include "control.hpp"
#include "sample/atoms.hpp"
#include "sample/atomselection.hpp"
#include "sample/center_of_mass.hpp"
#include "sample/coordinate_set.hpp"
#include "sample/frame.hpp"

namespace {
void fill(Frame &fr, const double *xyz, size_t n) {
    fr.number_of_atoms = n;
    for (size_t i = 0; i < n; i++) {
        fr.x.push_back(xyz[3 * i]);
        fr.y.push_back(xyz[3 * i + 1]);
        fr.z.push_back(xyz[3 * i + 2]);
    }
}
}  // namespace

extern "C" {
// mass of the atoms that carry database label `label` (the label must have been registered: ref_sample_name_reg)
void ref_mass_reg(const char *label, double mass) { Database::Inst()->masses.reg(Database::Inst()->atomIDs.get(label), mass); }

// Fit(atoms, cs, all, manip, cs_ref, ref): xyz [natoms][3] is the frame to fit (whole system), xyz_ref [natoms][3] the frame
// that holds the reference structure; sel_ref (ascending) selects the atoms the fit is computed on in both, sel_manip
// (ascending) the atoms that are moved.  mode 0: the fitted coordinates as Fit leaves them; out: double [natoms][3].
// com_out (may be NULL): CenterOfMass of the sel_ref atoms of xyz before the fit.
void ref_fit(const char *pdbfile, const double *xyz, const double *xyz_ref, size_t natoms, const size_t *sel_ref, size_t nref,
             const size_t *sel_manip, size_t nmanip, double *out, double *com_out) {
    Atoms atoms(pdbfile, "pdb");
    Frame fr, frr;
    fill(fr, xyz, natoms);
    fill(frr, xyz_ref, natoms);
    RangeAtomselection all(0, natoms - 1);
    IndexAtomselection ref(std::vector<size_t>(sel_ref, sel_ref + nref));
    IndexAtomselection manip(std::vector<size_t>(sel_manip, sel_manip + nmanip));
    CartesianCoordinateSet cs(fr, &all);
    CartesianCoordinateSet cs_ref(frr, &ref);
    if (com_out) {
        CartesianCoor3D c = CenterOfMass(atoms, cs, &all, &ref);
        com_out[0] = c.x;
        com_out[1] = c.y;
        com_out[2] = c.z;
    }
    Fit(atoms, cs, &all, &manip, cs_ref, &ref);
    for (size_t i = 0; i < natoms; i++) {
        out[3 * i] = cs.c1[i];
        out[3 * i + 1] = cs.c2[i];
        out[3 * i + 2] = cs.c3[i];
    }
}
}
