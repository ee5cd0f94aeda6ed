// ref_sample_wrap.cpp — the REFERENCE's own structure and selection readers (src/sample/atoms.cpp: PDB ATOM names -> database
// labels -> atom IDs; src/sample/atomselection_reader.cpp: ndx groups, PDB beta / segid selections; src/sample/atomselection.cpp:
// index / range selections) with its database.cpp, compiled where they lie into oracle/_ref/libparams_ref.so.  The database's
// name patterns are registered through its own reg() (its XML reader needs libxml2).  Test infrastructure: pins the product's
// control plane (csrc/host/control.cpp).
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "control.hpp"
#include "sample/atoms.hpp"
#include "sample/atomselection.hpp"
#include "sample/atomselection_reader.hpp"

namespace {
size_t copy_out(IAtomselection *s, size_t *out, size_t cap) {
    size_t n = s->size();
    for (size_t i = 0; i < n && i < cap && out; i++) out[i] = (*s)[i];
    return n;
}
}  // namespace

extern "C" {
void ref_sample_name_reg(const char *label, const char *regexp) {
    Database::Inst()->names.pdb.reg(label, regexp);
    Database::Inst()->atomIDs.reg(label);
}
// Atoms::add: the database label of every ATOM record, written as '\n'-separated text; returns the number of atoms
size_t ref_atoms_labels(const char *pdbfile, char *out, size_t cap) {
    Atoms atoms(pdbfile, "pdb");
    std::string all;
    for (size_t i = 0; i < atoms.size(); i++) all += Database::Inst()->atomIDs.rget(atoms[i]) + "\n";
    if (out && cap) {
        std::strncpy(out, all.c_str(), cap - 1);
        out[cap - 1] = 0;
    }
    return atoms.size();
}
size_t ref_select_pdb(const char *file, const char *selector, const char *expression, size_t *out, size_t cap) {
    IAtomselection *s = AtomselectionReader::read_pdb(file, selector, expression);
    size_t n = copy_out(s, out, cap);
    delete s;
    return n;
}
// one group of an ndx file by name (after the expression filter); returns its size, or (size_t)-1 when the group is absent
size_t ref_select_ndx(const char *file, const char *selector, const char *expression, const char *group, size_t *out, size_t cap) {
    std::map<std::string, IAtomselection *> m = AtomselectionReader::read_ndx(file, selector, expression);
    size_t n = (size_t)-1;
    if (m.find(group) != m.end()) n = copy_out(m[group], out, cap);
    for (std::map<std::string, IAtomselection *>::iterator i = m.begin(); i != m.end(); ++i) delete i->second;
    return n;
}
size_t ref_select_range(size_t from, size_t to, size_t *out, size_t cap) {
    RangeAtomselection s(from, to);
    return copy_out(&s, out, cap);
}
}
