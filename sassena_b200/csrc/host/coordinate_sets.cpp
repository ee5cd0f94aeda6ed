// coordinate_sets.cpp — what the reference does to a frame between reading it and staging it ("next" row):
//   motion walkers          reference src/sample/motion_walker.cpp:30-400
//   centre of mass / fit    reference src/sample/center_of_mass.cpp:60-275
//   CoordinateSets::load    reference src/sample/coordinate_sets.cpp:58-243 (set-up), :245-353 (per frame)
// All arithmetic is double on the whole system (CartesianCoordinateSet keeps coor2_t = double, coordinate_set.cpp:75-96);
// the narrowing to float happens in the stager.  Selections are walked the way the reference walks them (two sorted
// index lists merged), so a selection that is not sorted ascending skips atoms exactly as it does there.
#include <algorithm>
#include <cmath>
#include <fstream>

#include "control.hpp"

namespace sassena {

namespace {

void identity4(double T[16]) {
    for (int i = 0; i < 16; i++) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
}
void translation4(const CartesianCoor3D &t, double T[16]) {  // T(3,0..2) = t: row vector (x,y,z,1)*T = r + t
    identity4(T);
    T[12] = t.x;
    T[13] = t.y;
    T[14] = t.z;
}
void matmul4(const double A[16], const double B[16], double C[16]) {  // ublas::prod(A, B)
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double s = 0;
            for (int k = 0; k < 4; k++) s += A[4 * i + k] * B[4 * k + j];
            C[4 * i + j] = s;
        }
}

// translation(timepos) = timepos * (displace*sampling*direction/|direction|)          motion_walker.cpp:357-363
struct LinearMotionWalker : MotionWalker {
    CartesianCoor3D t;
    LinearMotionWalker(double displace, long sampling, CartesianCoor3D dir) { t = ((displace * sampling) * dir) / dir.length(); }
    void transform(size_t timepos, double T[16]) override { translation4((double)timepos * t, T); }
};
// translation = displace*direction/|direction|                                       motion_walker.cpp:379-385
struct FixedMotionWalker : MotionWalker {
    CartesianCoor3D t;
    FixedMotionWalker(double displace, CartesianCoor3D dir) { t = (displace * dir) / dir.length(); }
    void transform(size_t, double T[16]) override { translation4(t, T); }
};
// translation = m_translate * sin(2 pi timepos frequency sampling)                   motion_walker.cpp:333-341
struct OscillationMotionWalker : MotionWalker {
    CartesianCoor3D t;
    double frequency;
    long sampling;
    OscillationMotionWalker(double displace, double f, long s, CartesianCoor3D dir) : frequency(f), sampling(s) {
        t = (displace * dir) / dir.length();
    }
    void transform(size_t timepos, double T[16]) override {
        translation4(std::sin(2 * M_PI * timepos * frequency * sampling) * t, T);
    }
};
// cumulative steps of fixed length in uniformly random directions                    motion_walker.cpp:282-331
struct RandomMotionWalker : MotionWalker {
    UniformOnSphere sphere;
    long sampling;
    double displace;
    std::vector<CartesianCoor3D> translations;
    RandomMotionWalker(double d, unsigned long seed, long s) : sphere((uint32_t)seed, 3), sampling(s), displace(d) {}
    void transform(size_t timepos, double T[16]) override {
        while (translations.size() <= timepos) {
            CartesianCoor3D old = translations.empty() ? CartesianCoor3D(0, 0, 0) : translations.back();
            std::vector<double> v = sphere();
            for (long i = 0; i < sampling - 1; i++) sphere();
            translations.push_back(old + displace * CartesianCoor3D(v[0], v[1], v[2]));
        }
        translation4(translations[timepos], T);
    }
};
// cumulative steps of normal length (seed) in uniformly random directions (seed+1)   motion_walker.cpp:118-190
// radius > 0: steps that would leave the sphere of that radius are redrawn          motion_walker.cpp:192-280
struct BrownianMotionWalker : MotionWalker {
    BoostNormal normal;
    UniformOnSphere sphere;
    long sampling;
    double displace, radius;
    std::vector<CartesianCoor3D> translations;
    BrownianMotionWalker(double d, unsigned long seed, long s, double r)
        : normal((uint32_t)seed), sphere((uint32_t)(seed + 1), 3), sampling(s), displace(d), radius(r) {}
    void transform(size_t timepos, double T[16]) override {
        while (translations.size() <= timepos) {
            CartesianCoor3D old = translations.empty() ? CartesianCoor3D(0, 0, 0) : translations.back();
            CartesianCoor3D nt;
            while (true) {
                double nr = normal();
                std::vector<double> v = sphere();
                for (long i = 0; i < sampling - 1; i++) {
                    normal();
                    sphere();
                }
                nt = old + (displace * nr) * CartesianCoor3D(v[0], v[1], v[2]);
                if (radius > 0 && nt.length() > radius) continue;
                break;
            }
            translations.push_back(nt);
        }
        translation4(translations[timepos], T);
    }
};
// cumulative small rotations about z, y, x with normal angles (degrees)             motion_walker.cpp:30-116
struct RotationalBrownianMotionWalker : MotionWalker {
    BoostNormal normal;
    long sampling;
    double displace;
    std::vector<std::vector<double>> transformations;
    RotationalBrownianMotionWalker(double d, unsigned long seed, long s) : normal((uint32_t)seed), sampling(s), displace(d) {}
    void transform(size_t timepos, double T[16]) override {
        while (transformations.size() <= timepos) {
            double old[16];
            if (transformations.empty())
                identity4(old);
            else
                std::copy(transformations.back().begin(), transformations.back().end(), old);
            double n1 = normal(), n2 = normal(), n3 = normal();
            for (long i = 0; i < sampling - 1; i++) normal();
            double a1 = displace * n1 * M_PI / 180, a2 = displace * n2 * M_PI / 180, a3 = displace * n3 * M_PI / 180;
            double r1[16], r2[16], r3[16], r12[16], nw[16], out[16];
            identity4(r1);
            identity4(r2);
            identity4(r3);
            r1[0] = std::cos(a1);  r1[1] = std::sin(a1);  r1[4] = -std::sin(a1);  r1[5] = std::cos(a1);
            r2[0] = std::cos(a2);  r2[2] = -std::sin(a2); r2[8] = std::sin(a2);   r2[10] = std::cos(a2);
            r3[5] = std::cos(a3);  r3[6] = std::sin(a3);  r3[9] = -std::sin(a3);  r3[10] = std::cos(a3);
            matmul4(r1, r2, r12);
            matmul4(r12, r3, nw);
            matmul4(old, nw, out);
            transformations.emplace_back(out, out + 16);
        }
        std::copy(transformations[timepos].begin(), transformations[timepos].end(), T);
    }
};

// walks two index lists the way the reference does (coordinate_set.cpp:150-170): `fn(i)` for every index present in both
template <typename F>
void merge_walk(size_t nsystem, const std::vector<size_t> &sub, F fn) {
    size_t ci = 0, si = 0;
    while (ci < nsystem && si < sub.size()) {
        const size_t s = sub[si];
        if (ci == s) {
            fn(ci);
            si++;
            ci++;
        } else if (ci > s) {
            si++;
        } else {
            ci++;
        }
    }
}

// largest-eigenvalue eigenvector of a symmetric 4x4 matrix (cyclic Jacobi)
void jacobi4(double A[4][4], double evec[4]) {
    double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int sweep = 0; sweep < 64; sweep++) {
        double off = 0;
        for (int p = 0; p < 4; p++)
            for (int q = p + 1; q < 4; q++) off += A[p][q] * A[p][q];
        if (off < 1e-300) break;
        for (int p = 0; p < 4; p++)
            for (int q = p + 1; q < 4; q++) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
                const double c = 1 / std::sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < 4; k++) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 4; k++) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 4; k++) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    int best = 0;
    for (int i = 1; i < 4; i++)
        if (A[i][i] > A[best][best]) best = i;
    for (int k = 0; k < 4; k++) evec[k] = V[k][best];
}

}  // namespace

MotionWalker *MotionWalker::create(const SampleMotionParameters &m) {  // coordinate_sets.cpp:120-148
    if (m.type == "linear") return new LinearMotionWalker(m.displace, m.sampling, m.direction);
    if (m.type == "fixed") return new FixedMotionWalker(m.displace, m.direction);
    if (m.type == "oscillation") return new OscillationMotionWalker(m.displace, m.frequency, m.sampling, m.direction);
    if (m.type == "randomwalk") return new RandomMotionWalker(m.displace, m.seed, m.sampling);
    if (m.type == "brownian") return new BrownianMotionWalker(m.displace, m.seed, m.sampling, 0.0);
    if (m.type == "rotationalbrownian") return new RotationalBrownianMotionWalker(m.displace, m.seed, m.sampling);
    if (m.type == "localbrownian") {
        if (m.displace > m.radius) throw Error("radius size for local brownian motion smaller than displacement!");
        return new BrownianMotionWalker(m.displace, m.seed, m.sampling, m.radius);
    }
    if (m.type == "none") return nullptr;
    throw Error("Motion type not understood");
}

const std::vector<size_t> &CoordinateSetsProcessor::sel(const std::string &name) const {
    auto it = s_.selections.find(name);
    if (it == s_.selections.end()) throw Error("selection not found: " + name);
    return it->second;
}

// CartesianCoordinateSet(frame, selection) of the reference structure (coordinate_sets.cpp:84-115,170-202)
std::vector<double> CoordinateSetsProcessor::reference_set(const SampleReferenceParameters &r, const Config &, size_t NF,
                                                           const std::function<void(size_t, float *)> &load_raw) const {
    const size_t natoms = s_.atom_ids.size();
    const std::vector<size_t> &rs = sel(r.selection);
    std::vector<float> raw(natoms * 3);
    if (r.type == "frame") {
        size_t fn = r.frame;
        if (fn >= NF) fn = NF - 1;  // "reference frame number in alignment larger than size of frameset. Setting to last frame!"
        load_raw(fn, raw.data());
    } else if (r.type == "file") {
        if (r.format != "pdb") throw Error("File format for alignment reference not understood, format=" + r.format);
        // PDBFrameset(file).read_frame(frame): frames end at "END" lines (frames.cpp:442-577)
        std::ifstream in(r.filepath.c_str());
        if (in.fail()) throw Error("Couldn't open reference file: " + r.filepath);
        std::vector<std::vector<float>> frames;
        std::vector<float> cur;
        std::string line;
        while (getline(in, line)) {
            if (line.compare(0, 6, "ATOM  ") == 0) {
                if (line.size() < 54) throw Error("short ATOM record in " + r.filepath);
                cur.push_back((float)atof(line.substr(30, 8).c_str()));
                cur.push_back((float)atof(line.substr(38, 8).c_str()));
                cur.push_back((float)atof(line.substr(46, 8).c_str()));
            } else if (line.compare(0, 3, "END") == 0) {
                if (!cur.empty()) frames.push_back(cur);
                cur.clear();
            }
        }
        if (!cur.empty()) frames.push_back(cur);
        if (frames.empty()) throw Error("no frames in reference file: " + r.filepath);
        size_t fn = r.frame;
        if (fn >= frames.size()) fn = frames.size() - 1;
        raw = frames[fn];
    } else {
        throw Error("Reference type not understood:" + r.type);
    }
    std::vector<double> out;
    out.reserve(rs.size() * 3);
    for (size_t idx : rs) {
        if (3 * idx + 2 >= raw.size()) throw Error("Atom Index out of bounds for frame! Does the structure file match the frames?");
        for (int c = 0; c < 3; c++) out.push_back(raw[3 * idx + c]);
    }
    return out;
}

CoordinateSetsProcessor::CoordinateSetsProcessor(const Config &cfg, const Database &db, const LoadedSample &s, size_t NF,
                                                 const std::function<void(size_t, float *)> &load_raw)
    : db_(db), s_(s) {
    for (auto &m : cfg.motions) {  // coordinate_sets.cpp:72-156
        Motion mw;
        mw.selection = m.selection;
        mw.reference_selection = m.reference.selection;
        sel(mw.selection);
        sel(mw.reference_selection);
        if (m.reference.type == "instant") {
            mw.has_reference = false;
        } else {
            mw.ref = reference_set(m.reference, cfg, NF, load_raw);
            mw.has_reference = true;
        }
        mw.walker.reset(MotionWalker::create(m));
        if (mw.walker) motions_.push_back(std::move(mw));
    }
    for (auto &al : cfg.alignments) {  // coordinate_sets.cpp:158-241
        Alignment a;
        a.type = al.type;
        a.selection = al.selection;
        a.reference_selection = al.reference.selection;
        sel(a.selection);
        sel(a.reference_selection);
        if (al.reference.type == "instant") {
            a.has_reference = false;
        } else {
            a.ref = reference_set(al.reference, cfg, NF, load_raw);
            a.has_reference = true;
        }
        if (a.type != "center" && a.type != "fittrans" && a.type != "fitrottrans" && a.type != "fitrot")
            throw Error("Fitting routine not understood: " + a.type + ". Use either of: center , fittrans , fitrottrans, fitrot");
        if (al.order == "pre")
            pre_.push_back(a);
        else if (al.order == "post")
            post_.push_back(a);
        else
            throw Error("Ordering of alignment not understood. Must be pre or post.");
    }
}

namespace {
struct Com {
    double x = 0, y = 0, z = 0;
};
}  // namespace

void CoordinateSetsProcessor::align(const Alignment &a, std::vector<double> &xyz) const {
    const size_t natoms = s_.atom_ids.size();
    auto mass = [&](size_t atom) { return db_.mass(s_.atom_ids[atom]); };
    // CenterOfMass(atoms, cset, system, selection): mass-weighted mean over the atoms of `selection` (center_of_mass.cpp:164-213)
    auto com_system = [&](const std::vector<size_t> &selection) {
        Com c;
        if (selection.empty() || natoms == 0) return c;
        double m = 0, xt = 0, yt = 0, zt = 0;
        merge_walk(natoms, selection, [&](size_t i) {
            const double mi = mass(i);
            m += mi;
            xt += xyz[3 * i] * mi;
            yt += xyz[3 * i + 1] * mi;
            zt += xyz[3 * i + 2] * mi;
        });
        c.x = xt / m;
        c.y = yt / m;
        c.z = zt / m;
        return c;
    };
    // CenterOfMass(atoms, cset, selection) of a set that holds exactly the atoms of `selection` (center_of_mass.cpp:245-268)
    auto com_set = [&](const std::vector<double> &set, const std::vector<size_t> &selection) {
        Com c;
        if (selection.empty()) return c;
        double m = 0, xt = 0, yt = 0, zt = 0;
        for (size_t i = 0; i < selection.size(); i++) {
            const double mi = mass(selection[i]);
            m += mi;
            xt += set[3 * i] * mi;
            yt += set[3 * i + 1] * mi;
            zt += set[3 * i + 2] * mi;
        }
        c.x = xt / m;
        c.y = yt / m;
        c.z = zt / m;
        return c;
    };
    auto translate = [&](double tx, double ty, double tz, const std::vector<size_t> &selection) {
        merge_walk(natoms, selection, [&](size_t i) {
            xyz[3 * i] += tx;
            xyz[3 * i + 1] += ty;
            xyz[3 * i + 2] += tz;
        });
    };
    // the reference structure and its selection: the stored reference, or (instant) the current frame with the whole system
    std::vector<size_t> system_sel;
    const std::vector<size_t> *refsel = &sel(a.reference_selection);
    const std::vector<double> *refset = &a.ref;
    std::vector<double> current;
    if (!a.has_reference) {
        system_sel.resize(natoms);
        for (size_t i = 0; i < natoms; i++) system_sel[i] = i;
        refsel = &system_sel;
        current = xyz;
        refset = &current;
    }
    // Fit (center_of_mass.cpp:60-160): mass-weighted least-squares rotation of the atoms of the reference selection onto the
    // reference structure about the centres of mass, then the fitted coordinates are written back for the atoms that are
    // in both the manipulated and the reference selection.  The reference takes the rotation from LAPACK dgesvd (Kabsch,
    // with the determinant correction); the same optimal proper rotation is obtained here from Horn's quaternion form.
    // (The reference's write-back indexes its reduced copy by ATOM index, center_of_mass.cpp:146-149, which is only in
    // bounds and correct when the reference selection is a prefix of the system; the intended element is used here.)
    auto fit = [&]() {
        const std::vector<size_t> &rs = *refsel;
        if (rs.size() * 3 != refset->size())
            throw Error("Fitting requires the reference and the target to contain the same number of atoms");
        std::vector<double> red;  // CoordinateSet(cs, system, ref_selection)
        std::vector<size_t> red_atoms;
        merge_walk(natoms, rs, [&](size_t i) {
            red.push_back(xyz[3 * i]);
            red.push_back(xyz[3 * i + 1]);
            red.push_back(xyz[3 * i + 2]);
            red_atoms.push_back(i);
        });
        if (red.size() != refset->size())
            throw Error("Fitting requires the reference and the target to contain the same number of atoms");
        const Com ref = com_set(*refset, rs), pos = com_set(red, rs);
        double K[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};  // sum m u v^T, u = reference, v = current
        for (size_t i = 0; i < rs.size(); i++) {
            const double mi = mass(rs[i]);
            const double v[3] = {red[3 * i] - pos.x, red[3 * i + 1] - pos.y, red[3 * i + 2] - pos.z};
            const double u[3] = {(*refset)[3 * i] - ref.x, (*refset)[3 * i + 1] - ref.y, (*refset)[3 * i + 2] - ref.z};
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) K[r][c] += mi * u[r] * v[c];
        }
        // maximise sum m u.(R v): quaternion matrix of S = sum m v u^T = K^T
        const double Sxx = K[0][0], Sxy = K[1][0], Sxz = K[2][0];
        const double Syx = K[0][1], Syy = K[1][1], Syz = K[2][1];
        const double Szx = K[0][2], Szy = K[1][2], Szz = K[2][2];
        double Nm[4][4] = {{Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx},
                           {Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz},
                           {Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy},
                           {Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz}};
        double qn[4];
        jacobi4(Nm, qn);
        const double n = std::sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
        const double w = qn[0] / n, x = qn[1] / n, y = qn[2] / n, z = qn[3] / n;
        const double R[3][3] = {{1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)},
                                {2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)},
                                {2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)}};
        for (size_t i = 0; i < rs.size(); i++) {
            const double b[3] = {red[3 * i] - pos.x, red[3 * i + 1] - pos.y, red[3 * i + 2] - pos.z};
            red[3 * i] = R[0][0] * b[0] + R[0][1] * b[1] + R[0][2] * b[2] + ref.x;
            red[3 * i + 1] = R[1][0] * b[0] + R[1][1] * b[1] + R[1][2] * b[2] + ref.y;
            red[3 * i + 2] = R[2][0] * b[0] + R[2][1] * b[1] + R[2][2] * b[2] + ref.z;
        }
        const std::vector<size_t> &manip = sel(a.selection);
        size_t ci = 0, si = 0;
        while (ci < manip.size() && si < red_atoms.size()) {
            if (manip[ci] == red_atoms[si]) {
                xyz[3 * manip[ci]] = red[3 * si];
                xyz[3 * manip[ci] + 1] = red[3 * si + 1];
                xyz[3 * manip[ci] + 2] = red[3 * si + 2];
                si++;
                ci++;
            } else if (manip[ci] > red_atoms[si]) {
                si++;
            } else {
                ci++;
            }
        }
    };
    if (a.type == "center") {  // coordinate_sets.cpp:272-275
        const Com o = com_system(sel(a.reference_selection));
        translate(-1.0 * o.x, -1.0 * o.y, -1.0 * o.z, sel(a.selection));
    } else if (a.type == "fittrans") {  // :276-279
        const Com ref = com_set(*refset, *refsel);
        const Com pos = com_system(sel(a.selection));
        translate(ref.x - pos.x, ref.y - pos.y, ref.z - pos.z, sel(a.selection));
    } else if (a.type == "fitrottrans") {  // :280-281
        fit();
    } else if (a.type == "fitrot") {  // :282-286: the fit moves the centre of mass; the reference adds the old one back
        const Com pos = com_system(sel(a.selection));
        fit();
        translate(pos.x, pos.y, pos.z, sel(a.selection));
    }
}

void CoordinateSetsProcessor::apply(size_t framenumber, std::vector<double> &xyz) const {
    const size_t natoms = s_.atom_ids.size();
    for (auto &a : pre_) align(a, xyz);
    for (auto &mw : motions_) {  // coordinate_sets.cpp:294-310
        double rx = 0, ry = 0, rz = 0;
        const std::vector<size_t> &rs = sel(mw.reference_selection);
        double m = 0, xt = 0, yt = 0, zt = 0;
        if (mw.has_reference) {
            for (size_t i = 0; i < rs.size(); i++) {
                const double mi = db_.mass(s_.atom_ids[rs[i]]);
                m += mi;
                xt += mw.ref[3 * i] * mi;
                yt += mw.ref[3 * i + 1] * mi;
                zt += mw.ref[3 * i + 2] * mi;
            }
        } else {
            merge_walk(natoms, rs, [&](size_t i) {
                const double mi = db_.mass(s_.atom_ids[i]);
                m += mi;
                xt += xyz[3 * i] * mi;
                yt += xyz[3 * i + 1] * mi;
                zt += xyz[3 * i + 2] * mi;
            });
        }
        if (!rs.empty() && natoms) {
            rx = xt / m;
            ry = yt / m;
            rz = zt / m;
        }
        double T[16];
        mw.walker->transform(framenumber, T);
        const double nx = -1.0 * rx, ny = -1.0 * ry, nz = -1.0 * rz;
        merge_walk(natoms, sel(mw.selection), [&](size_t i) {
            double x = xyz[3 * i] + nx, y = xyz[3 * i + 1] + ny, z = xyz[3 * i + 2] + nz;
            // newpos = prod(pos, T), pos = (x, y, z, 1)
            const double px = ((x * T[0] + y * T[4]) + z * T[8]) + 1.0 * T[12];
            const double py = ((x * T[1] + y * T[5]) + z * T[9]) + 1.0 * T[13];
            const double pz = ((x * T[2] + y * T[6]) + z * T[10]) + 1.0 * T[14];
            xyz[3 * i] = px + rx;
            xyz[3 * i + 1] = py + ry;
            xyz[3 * i + 2] = pz + rz;
        });
    }
    for (auto &a : post_) align(a, xyz);
}

}  // namespace sassena
