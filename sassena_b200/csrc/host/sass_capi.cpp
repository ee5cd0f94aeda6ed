// sass_capi.cpp — C wrapper of the host layer (include/sassena_host.h).  Exceptions stop here.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "control.hpp"
#include "dcd.hpp"
#include "xdr_traj.hpp"
#include "h5mini.hpp"
#include "sassena_host.hpp"

using namespace sassena;

struct sass_xdr {
    std::unique_ptr<sassena::XdrFrameset> fs;
};
struct sass_dcd {
    sassena::DCDFrameset fs;
    explicit sass_dcd(const std::string &p) : fs(p) {}
};

struct sass_params {
    Params params;
};

struct sass_job {
    sassena::Job job;
    sass_params params;  // copy of the Params slice of job.cfg for sass_job_params
};

namespace {
thread_local std::string g_err;

template <typename F>
int guard(F f) {
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    } catch (...) {
        g_err = "unknown error";
        return 1;
    }
}

bool parse_bool(const std::string &v) {
    // xml_interface.cpp:56-61: true/false in any case, or 1/0
    std::string s;
    for (char c : v) s += (char)tolower(c);
    if (s == "true" || s == "1") return true;
    if (s == "false" || s == "0") return false;
    throw Error("boolean value not understood: " + v);
}

class FnSink : public IResultSink {
    sass_write_fn fn_;
    void *user_;

   public:
    FnSink(sass_write_fn fn, void *user) : fn_(fn), user_(user) {}
    void write(CartesianCoor3D q, const double *fqt, size_t NF, std::complex<double> fq, std::complex<double> fq2) override {
        if (!fn_) return;
        double qq[3] = {q.x, q.y, q.z};
        double a[2] = {fq.real(), fq.imag()};
        double b[2] = {fq2.real(), fq2.imag()};
        fn_(user_, qq, fqt, NF, a, b);
    }
};
}  // namespace

extern "C" {

const char *sass_last_error(void) { return g_err.c_str(); }

sass_params *sass_params_new(void) { return new sass_params(); }
void sass_params_free(sass_params *p) { delete p; }

int sass_params_set(sass_params *p, const char *key, const char *value) {
    return guard([&] {
        if (!p || !key || !value) throw Error("sass_params_set: NULL argument");
        const std::string k(key), v(value);
        ScatteringParameters &s = p->params.scattering;
        const std::string o = "scattering.average.orientation.";
        if (k == "scattering.type") s.type = v;
        else if (k == "scattering.dsp.type") s.dsp_type = v;
        else if (k == "scattering.dsp.method") s.dsp_method = v;
        else if (k == o + "type") s.orientation_type = v;
        else if (k == o + "axis.x") s.axis.x = atof(value);
        else if (k == o + "axis.y") s.axis.y = atof(value);
        else if (k == o + "axis.z") s.axis.z = atof(value);
        else if (k == o + "vectors.type") s.vectors.type = v;
        else if (k == o + "vectors.algorithm") s.vectors.algorithm = v;
        else if (k == o + "vectors.resolution") s.vectors.resolution = strtoull(value, nullptr, 10);
        else if (k == o + "vectors.seed") s.vectors.seed = (uint32_t)strtoul(value, nullptr, 10);
        else if (k == o + "multipole.type") s.multipole.type = v;
        else if (k == o + "multipole.moments.type") s.multipole.moments_type = v;
        else if (k == o + "multipole.moments.resolution") s.multipole.resolution = atol(value);
        else if (k == "limits.stage.memory.data") p->params.limits.stage_memory_data = strtoull(value, nullptr, 10);
        else if (k == "limits.stage.stream") p->params.limits.stage_stream = parse_bool(v);
        else if (k == "stager.dump") p->params.stager.dump = parse_bool(v);
        else if (k == "stager.file") p->params.stager.file = p->params.stager.filepath = v;
        else if (k == "stager.format") p->params.stager.format = v;
        else if (k == "limits.decomposition.utilization") p->params.limits.decomposition.utilization = atof(value);
        else if (k == "limits.decomposition.partitions.automatic")
            p->params.limits.decomposition.partitions_automatic = parse_bool(v);
        else if (k == "limits.computation.scan_snap")
            p->params.limits.coherent_scan_snap = parse_bool(v);
        else if (k == "limits.computation.scan")
            p->params.limits.coherent_scan = strtoull(value, nullptr, 10);
        else if (k == "limits.decomposition.coherent")
            p->params.limits.coherent_sharding = value;
        else if (k == "limits.decomposition.partitions.size")
            p->params.limits.decomposition.partitions_size = strtoull(value, nullptr, 10);
        else if (k == "scattering.target")
            throw Error("scattering.target is obsolete. Use stager.target instead.");  // parameters.cpp:404-407
        else throw Error("unknown parameter key: " + k);
    });
}

int sass_params_set_vectors(sass_params *p, const double *xyz, size_t n) {
    return guard([&] {
        if (!p || (!xyz && n)) throw Error("sass_params_set_vectors: NULL argument");
        auto &v = p->params.scattering.vectors.vectors;
        v.clear();
        for (size_t i = 0; i < n; i++) v.push_back(CartesianCoor3D(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
    });
}

int sass_params_set_moments(sass_params *p, const long *lm, size_t n) {
    return guard([&] {
        if (!p || (!lm && n)) throw Error("sass_params_set_moments: NULL argument");
        auto &m = p->params.scattering.multipole.moments;
        m.clear();
        for (size_t i = 0; i < n; i++) m.push_back(std::make_pair(lm[2 * i], lm[2 * i + 1]));
    });
}

int sass_params_create(sass_params *p) {
    return guard([&] {
        if (!p) throw Error("sass_params_create: NULL argument");
        ScatteringParameters &s = p->params.scattering;
        if (s.orientation_type == "vectors") s.vectors.create();
        else if (s.orientation_type == "multipole") s.multipole.create();
        else if (s.orientation_type != "none")
            throw Error("scattering.average.orientation.type not understood: " + s.orientation_type);
    });
}

size_t sass_params_num_vectors(const sass_params *p) { return p ? p->params.scattering.vectors.vectors.size() : 0; }
int sass_params_get_vectors(const sass_params *p, double *out) {
    return guard([&] {
        if (!p || !out) throw Error("NULL argument");
        size_t i = 0;
        for (auto &v : p->params.scattering.vectors.vectors) {
            out[3 * i] = v.x;
            out[3 * i + 1] = v.y;
            out[3 * i + 2] = v.z;
            i++;
        }
    });
}
size_t sass_params_num_moments(const sass_params *p) { return p ? p->params.scattering.multipole.moments.size() : 0; }
int sass_params_get_moments(const sass_params *p, long *out) {
    return guard([&] {
        if (!p || !out) throw Error("NULL argument");
        size_t i = 0;
        for (auto &m : p->params.scattering.multipole.moments) {
            out[2 * i] = m.first;
            out[2 * i + 1] = m.second;
            i++;
        }
    });
}

int sass_scatter_run(const sass_params *p, const sass_comm_vtbl *comm, const sass_backend_vtbl *backend, sgpu_ctx *ctx,
                     size_t NA, size_t NF, const float *frames, const double *b_const, sass_factors_fn ffn,
                     void *fuser, const double *qvectors, size_t NQ, sass_write_fn wfn, void *wuser, int *has_device,
                     char *timers, size_t timers_cap) {
    return guard([&] {
        if (!p || !frames || !qvectors) throw Error("sass_scatter_run: NULL argument");
        if (!b_const && !ffn) throw Error("sass_scatter_run: no scattering factors given");
        std::shared_ptr<ICommunicator> c;
        if (comm) c = std::make_shared<CallbackCommunicator>(*comm, false);
        else c = std::make_shared<SingleCommunicator>();
        Sample sample;
        sample.NA = NA;
        sample.NF = NF;
        sample.frames = frames;
        if (ffn) sample.factors = [=](double ql, double *b) { ffn(fuser, ql, b, NA); };
        else sample.factors = [=](double, double *b) { memcpy(b, b_const, NA * sizeof(double)); };
        std::vector<CartesianCoor3D> qv;
        for (size_t i = 0; i < NQ; i++) qv.push_back(CartesianCoor3D(qvectors[3 * i], qvectors[3 * i + 1], qvectors[3 * i + 2]));
        FnSink sink(wfn, wuser);
        const SgpuBackend &be = backend ? *backend : default_backend();
        std::unique_ptr<IScatterDevice> dev(ScatterDeviceFactory::create(c, sample, &sink, qv, p->params, be, ctx));
        if (has_device) *has_device = dev ? 1 : 0;
        if (timers && timers_cap) timers[0] = 0;
        if (!dev) return;
        dev->run();
        if (timers && timers_cap) {
            std::string t;
            Timer &tm = dev->getTimer();
            for (auto &k : tm.keys()) {
                char buf[160];
                snprintf(buf, sizeof(buf), "%s=%.6f:%zu;", k.c_str(), tm.sum(k), tm.count(k));
                t += buf;
            }
            strncpy(timers, t.c_str(), timers_cap - 1);
            timers[timers_cap - 1] = 0;
        }
    });
}

int sass_motion_transforms(const char *type, double displace, double frequency, double radius, unsigned long seed, long sampling,
                           const double dir[3], size_t n, double *out) {
    return guard([&] {
        SampleMotionParameters m;
        m.type = type;
        m.displace = displace;
        m.frequency = frequency;
        m.radius = radius;
        m.seed = seed;
        m.sampling = sampling;
        m.direction = CartesianCoor3D(dir[0], dir[1], dir[2]);
        std::unique_ptr<MotionWalker> w(MotionWalker::create(m));
        if (!w) throw Error("Motion type none has no walker");
        for (size_t t = 0; t < n; t++) w->transform(t, out + 16 * t);
    });
}

int sass_div_assignment(size_t NN, size_t rank, size_t NAF, size_t *offset, size_t *size, size_t *max) {
    return guard([&] {
        DivAssignment a(NN, rank, NAF);
        *offset = a.offset();
        *size = a.size();
        *max = a.max();
    });
}
int sass_mod_assignment(size_t NN, size_t rank, size_t NAF, size_t *offset, size_t *size, size_t *max) {
    return guard([&] {
        ModAssignment a(NN, rank, NAF);
        *offset = a.offset();
        *size = a.size();
        *max = a.max();
    });
}
int sass_decomposition_plan(size_t nn, size_t nq, size_t naf, size_t elbytesize, size_t nmaxbytesize, double utilization,
                            int automatic, size_t manual_size, size_t *partitions, size_t *partitionsize,
                            size_t *penalty) {
    return guard([&] {
        DecompositionLimits lim;
        lim.utilization = utilization;
        lim.partitions_automatic = automatic != 0;
        lim.partitions_size = manual_size;
        DecompositionPlan plan(nn, nq, naf, elbytesize, nmaxbytesize, lim);
        *partitions = plan.partitions();
        *partitionsize = plan.partitionsize();
        if (penalty) *penalty = plan.penalty();
    });
}
size_t sass_create_from_scans(const double *scans, size_t nscans, double *out, size_t cap) {
    size_t n = 0;
    guard([&] {
        std::vector<ScatteringVectorsScan> sc;
        for (size_t i = 0; i < nscans; i++) {
            ScatteringVectorsScan s;
            s.basevector = CartesianCoor3D(scans[7 * i], scans[7 * i + 1], scans[7 * i + 2]);
            s.from = scans[7 * i + 3];
            s.to = scans[7 * i + 4];
            s.points = (size_t)scans[7 * i + 5];
            s.exponent = scans[7 * i + 6];
            sc.push_back(s);
        }
        auto q = create_from_scans(sc);
        n = q.size();
        if (out)
            for (size_t i = 0; i < n && i < cap; i++) {
                out[3 * i] = q[i].x;
                out[3 * i + 1] = q[i].y;
                out[3 * i + 2] = q[i].z;
            }
    });
    return n;
}

namespace {
// exposes the protected init_subvectors for sass_init_subvectors
struct SubvectorProbe : AllVectorsScatterDevice {
    using AllVectorsScatterDevice::AllVectorsScatterDevice;
    using AbstractVectorsScatterDevice::init_subvectors;
};
}  // namespace

size_t sass_init_subvectors(const sass_params *p, const double q[3], double *out, size_t cap) {
    size_t n = 0;
    guard([&] {
        if (!p || !q) throw Error("NULL argument");
        // init_subvectors only reads params; build a device shell around a dummy backend-less context
        static sass_backend_vtbl none = {};
        Sample sample;
        sample.NA = 1;
        sample.NF = 1;
        auto c = std::make_shared<SingleCommunicator>();
        SubvectorProbe probe(c, c, sample, {}, 1, nullptr, p->params, none, reinterpret_cast<sgpu_ctx *>(&none));
        CartesianCoor3D qq(q[0], q[1], q[2]);
        probe.init_subvectors(qq);
        const auto &sv = probe.subvectors();
        n = sv.size();
        if (out)
            for (size_t i = 0; i < n && i < cap; i++) {
                out[3 * i] = sv[i].x;
                out[3 * i + 1] = sv[i].y;
                out[3 * i + 2] = sv[i].z;
            }
    });
    return n;
}

int sass_dcd_open(const char *path, size_t first, size_t last, int last_set, size_t stride, sass_dcd **out) {
    return guard([&] {
        if (!path || !out) throw Error("sass_dcd_open: NULL argument");
        std::unique_ptr<sass_dcd> d(new sass_dcd(path));
        d->fs.trim_index(first, last, last_set != 0, stride);
        *out = d.release();
    });
}
int sass_dcd_info(const sass_dcd *d, size_t *nframes, size_t *natoms, int *has_unitcell) {
    return guard([&] {
        if (!d) throw Error("sass_dcd_info: NULL argument");
        if (nframes) *nframes = d->fs.number_of_frames;
        if (natoms) *natoms = d->fs.number_of_atoms;
        if (has_unitcell) *has_unitcell = d->fs.has_unitcell() ? 1 : 0;
    });
}
int sass_dcd_read(sass_dcd *d, size_t first, size_t count, float *out) {
    return guard([&] {
        if (!d || !out) throw Error("sass_dcd_read: NULL argument");
        if (first + count > d->fs.number_of_frames) throw Error("sass_dcd_read: frame range out of bounds");
        d->fs.read_frames(first, count, out);
    });
}
void sass_dcd_close(sass_dcd *d) { delete d; }
int sass_dcd_write(const char *path, const float *xyz, size_t NF, size_t NA) {
    return guard([&] {
        if (!path || !xyz) throw Error("sass_dcd_write: NULL argument");
        DCDCoordinateWriter w(path, NF, NA);
        w.init();
        w.prepare();
        w.write(xyz, 0, NF);
    });
}

int sass_xdr_open(const char *path, const char *format, size_t first, size_t last, int last_set, size_t stride,
                  sass_xdr **out) {
    return guard([&] {
        if (!path || !format || !out) throw Error("sass_xdr_open: NULL argument");
        std::unique_ptr<sass_xdr> d(new sass_xdr());
        const std::string f(format);
        if (f == "xtc")
            d->fs.reset(new XTCFrameset(path));
        else if (f == "trr")
            d->fs.reset(new TRRFrameset(path));
        else
            throw Error("sass_xdr_open: format must be xtc or trr");
        d->fs->trim_index(first, last, last_set != 0, stride);
        *out = d.release();
    });
}
int sass_xdr_info(const sass_xdr *d, size_t *nframes, size_t *natoms) {
    return guard([&] {
        if (!d) throw Error("sass_xdr_info: NULL argument");
        if (nframes) *nframes = d->fs->number_of_frames;
        if (natoms) *natoms = d->fs->number_of_atoms;
    });
}
int sass_xdr_read(sass_xdr *d, size_t first, size_t count, float *out, double *box) {
    return guard([&] {
        if (!d || !out) throw Error("sass_xdr_read: NULL argument");
        if (first + count > d->fs->number_of_frames) throw Error("sass_xdr_read: frame range out of bounds");
        const size_t na = d->fs->number_of_atoms;
        for (size_t i = 0; i < count; i++) d->fs->read_frame(first + i, out + i * na * 3, box ? box + 9 * i : nullptr);
    });
}
void sass_xdr_close(sass_xdr *d) { delete d; }

static void h5_walk(const h5::Group &g, const std::string &prefix, sass_h5_dataset_fn fn, void *user) {
    for (auto &d : g.datasets) {
        const std::string p = prefix + d.name;
        const void *data = d.kind == h5::Dataset::F64 ? (const void *)d.f64.data() : (const void *)d.bytes.data();
        const size_t nbytes = d.kind == h5::Dataset::F64 ? d.f64.size() * sizeof(double) : d.bytes.size();
        fn(user, p.c_str(), (int)d.kind, d.dims.size(), d.dims.data(), d.maxdims.empty() ? nullptr : d.maxdims.data(),
           d.chunk.empty() ? nullptr : d.chunk.data(), data, nbytes);
    }
    for (auto &s : g.groups) h5_walk(s, prefix + s.name + "/", fn, user);
}

int sass_h5_read(const char *path, sass_h5_dataset_fn fn, void *user) {
    return guard([&] {
        if (!path || !fn) throw Error("sass_h5_read: NULL argument");
        h5_walk(h5::read_file(path), "", fn, user);
    });
}

int sass_h5_write_signal(const char *path, size_t NF, size_t chunksize, int flags, int resume, const char *rawconfig,
                         const char *config, const char *database, size_t n, const double *q, const double *fqt,
                         const double *fq, const double *fq2, size_t *rows_total) {
    return guard([&] {
        if (!path || (n && (!q || !fqt || !fq || !fq2))) throw Error("sass_h5_write_signal: NULL argument");
        SignalFileH5 f(path, NF, chunksize, flags & 1, flags & 2, flags & 4, flags & 8);
        f.set_meta(rawconfig ? rawconfig : "", config ? config : "", database ? database : "");
        if (resume) f.init();
        for (size_t i = 0; i < n; i++) f.write(q + 3 * i, fqt + i * 2 * NF, fq + 2 * i, fq2 + 2 * i);
        f.flush();
        if (rows_total) *rows_total = f.rows();
    });
}

}  // extern "C"

// ---- control plane ----
int sass_job_load(const char *config_file, sass_job **out) {
    return guard([&] {
        if (!config_file || !out) throw Error("sass_job_load: NULL argument");
        std::unique_ptr<sass_job> j(new sass_job);
        j->job.load(config_file);
        j->params.params = static_cast<const Params &>(j->job.cfg);
        *out = j.release();
    });
}

int sass_job_load_overwrite(const char *config_file, const char *const *keys, const char *const *values, size_t n,
                            sass_job **out) {
    return guard([&] {
        if (!config_file || !out || (n && (!keys || !values))) throw Error("sass_job_load_overwrite: NULL argument");
        std::vector<std::pair<std::string, std::string>> kv;
        for (size_t i = 0; i < n; i++) {
            if (!keys[i] || !values[i]) throw Error("sass_job_load_overwrite: NULL key or value");
            kv.emplace_back(keys[i], values[i]);
        }
        std::unique_ptr<sass_job> j(new sass_job);
        j->job.load(config_file, kv);
        j->params.params = static_cast<const Params &>(j->job.cfg);
        *out = j.release();
    });
}

void sass_job_free(sass_job *j) { delete j; }

int sass_job_info(const sass_job *j, size_t *natoms, size_t *ntarget, size_t *nframes, size_t *nqvectors) {
    return guard([&] {
        if (!j) throw Error("sass_job_info: NULL job");
        if (natoms) *natoms = j->job.sample.atom_ids.size();
        if (ntarget) *ntarget = j->job.sample.target.size();
        if (nframes) *nframes = j->job.sample.NF;
        if (nqvectors) *nqvectors = j->job.cfg.qvectors.size();
    });
}

int sass_job_qvectors(const sass_job *j, double *q_out) {
    return guard([&] {
        if (!j || !q_out) throw Error("sass_job_qvectors: NULL argument");
        size_t i = 0;
        for (auto &q : j->job.cfg.qvectors) {
            q_out[i++] = q.x;
            q_out[i++] = q.y;
            q_out[i++] = q.z;
        }
    });
}

int sass_job_factors(const sass_job *j, double ql, double *b_out) {
    return guard([&] {
        if (!j || !b_out) throw Error("sass_job_factors: NULL argument");
        j->job.factors->update(ql, b_out);
    });
}

int sass_job_frames(const sass_job *j, const float **frames) {
    return guard([&] {
        if (!j || !frames) throw Error("sass_job_frames: NULL argument");
        *frames = j->job.sample.frames.data();
    });
}

int sass_job_selection(const sass_job *j, const char *name, size_t *ids, size_t cap, size_t *n) {
    return guard([&] {
        if (!j || !name || !n) throw Error("sass_job_selection: NULL argument");
        auto it = j->job.sample.selections.find(name);
        if (it == j->job.sample.selections.end()) throw Error(std::string("selection not found: ") + name);
        *n = it->second.size();
        for (size_t i = 0; i < it->second.size() && i < cap && ids; i++) ids[i] = it->second[i];
    });
}

const sass_params *sass_job_params(const sass_job *j) { return j ? &j->params : nullptr; }
const char *sass_job_signal_file(const sass_job *j) { return j ? j->job.cfg.signal_filepath.c_str() : nullptr; }
const char *sass_job_option(const sass_job *j, const char *key) {
    if (!j || !key) return nullptr;
    const Config &c = j->job.cfg;
    const std::string k(key);
    if (k == "sample.structure.file") return c.structure_filepath.c_str();
    if (k == "sample.structure.format") return c.structure_format.c_str();
    if (k == "stager.target") return c.stager_target.c_str();
    if (k == "stager.dump") return c.stager.dump ? "true" : "false";
    if (k == "stager.file") return c.stager.filepath.c_str();
    if (k == "stager.format") return c.stager.format.c_str();
    if (k == "scattering.signal.file") return c.signal_filepath.c_str();
    return nullptr;
}

int sass_job_stage(sass_job *j, const sass_comm_vtbl *comm, const sass_backend_vtbl *backend, sgpu_ctx *ctx, size_t *staged_bytes,
                   char *report, size_t report_cap) {
    return guard([&] {
        if (!j) throw Error("sass_job_stage: NULL argument");
        std::shared_ptr<ICommunicator> c;
        if (comm) c = std::make_shared<CallbackCommunicator>(*comm, false);
        else c = std::make_shared<SingleCommunicator>();
        const SgpuBackend &be = backend ? *backend : default_backend();
        std::string rep;
        size_t n = j->job.stage(c, be, ctx, &rep);
        if (staged_bytes) *staged_bytes = n;
        if (report && report_cap) {
            strncpy(report, rep.c_str(), report_cap - 1);
            report[report_cap - 1] = 0;
        }
    });
}

int sass_job_run(sass_job *j, const char *signal_dir, const sass_comm_vtbl *comm, const sass_backend_vtbl *backend,
                 sgpu_ctx *ctx, size_t *written, char *report, size_t report_cap) {
    return guard([&] {
        if (!j || !signal_dir) throw Error("sass_job_run: NULL argument");
        std::shared_ptr<ICommunicator> c;
        if (comm) c = std::make_shared<CallbackCommunicator>(*comm, false);
        else c = std::make_shared<SingleCommunicator>();
        const SgpuBackend &be = backend ? *backend : default_backend();
        std::string rep;
        size_t n = j->job.run(signal_dir, c, be, ctx, &rep);
        if (written) *written = n;
        if (report && report_cap) {
            strncpy(report, rep.c_str(), report_cap - 1);
            report[report_cap - 1] = 0;
        }
    });
}
