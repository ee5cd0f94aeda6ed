// ref_db_wrap.cpp — evaluates the REFERENCE's own Database tables (src/control/database.cpp compiled where it lies over the
// shims: volumes, exclusion factors, scattering factors with their powf / sqrt overload roundings, database.cpp:391-528, and the
// PDB atom-name regular expressions, :309-340).  The XML reader is compiled but not used (libxml2 is absent): the tables are
// registered through the reference's own reg() methods.  Test infrastructure: pins the product's ScatterFactors / Database
// (csrc/host/control.cpp) and the restatement in tests/test_control_plane.py.
#include <cstring>
#include <string>
#include <vector>

#include "control/database.hpp"

extern "C" {
// table: 0 = sizes (volumes), 1 = exclusionfactors, 2 = scatterfactors
void ref_db_reg(int table, size_t ID, const double *constants, size_t n, size_t function_type) {
    std::vector<double> c(constants, constants + n);
    Database *db = Database::Inst();
    if (table == 0) db->volumes.reg(ID, c, function_type);
    else if (table == 1) db->exclusionfactors.reg(ID, c, function_type);
    else db->sfactors.reg(ID, c, function_type);
}
double ref_db_volume(size_t ID) { return Database::Inst()->volumes.get(ID); }
double ref_db_exclusion(size_t ID, double effvolume, double q) { return Database::Inst()->exclusionfactors.get(ID, effvolume, q); }
double ref_db_sfactor(size_t ID, double q) { return Database::Inst()->sfactors.get(ID, q); }
// ScatterFactors::update for one atom (scatter_factors.cpp:56-78): sf - background * efactor(kappa * volume)
double ref_db_effective(size_t ID, double q, double kappa, double background_sl) {
    Database *db = Database::Inst();
    double sf = db->sfactors.get(ID, q);
    double v = db->volumes.get(ID);
    double efactor = db->exclusionfactors.get(ID, kappa * v, q);
    return sf - background_sl * efactor;
}
void ref_db_name_reg(const char *label, const char *regexp) { Database::Inst()->names.pdb.reg(label, regexp); }
// resolves a PDB atom name to its element label.  Only for names the database knows: the reference rejects unknown / ambiguous
// names with a bare `throw;` (database.cpp:323,334), which terminates the process.
void ref_db_name_get(const char *testlabel, char *out, size_t cap) {
    std::string l = Database::Inst()->names.pdb.get(testlabel);
    std::strncpy(out, l.c_str(), cap - 1);
    out[cap - 1] = 0;
}
}
