// sassena_host.cpp — see sassena_host.hpp.  Reference citations are relative to benlabs/sassena v1.4.2.
#include "sassena_host.hpp"

#include "dcd.hpp"

#include <algorithm>
#include <chrono>
#include <thread>
#include <cmath>
#include <random>

namespace sassena {

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

// ---------------------------------------------------------------------------------------------------------------
// coor3d
// ---------------------------------------------------------------------------------------------------------------
double CartesianCoor3D::length() const { return std::sqrt(std::pow(x, 2) + std::pow(y, 2) + std::pow(z, 2)); }

CartesianVectorBase::CartesianVectorBase(CartesianCoor3D axis) {
    CartesianCoor3D ek(0, 0, 1), ej(0, 1, 0);
    CartesianCoor3D ez = axis / axis.length();
    CartesianCoor3D ekez = ek.cross_product(ez);
    if (ekez.length() == 0) ekez = ej.cross_product(ez);
    CartesianCoor3D er = ekez / ekez.length();
    CartesianCoor3D ezer = ez.cross_product(er);
    CartesianCoor3D ephi = ezer / ezer.length();
    base_.push_back(er);
    base_.push_back(ephi);
    base_.push_back(ez);
}

CartesianCoor3D CartesianVectorBase::project(CartesianCoor3D vec) {
    return CartesianCoor3D(vec * base_[0], vec * base_[1], vec * base_[2]);
}

// ---------------------------------------------------------------------------------------------------------------
// decomposition (assignment.cpp:27-132)
// ---------------------------------------------------------------------------------------------------------------
DivAssignment::DivAssignment(size_t NN, size_t rank, size_t NAF) : NN_(NN), rank_(rank), NAF_(NAF) {
    if (NN == 0) throw Error("DivAssignment: NN == 0");
    size_t first = (rank_ * NAF_) / NN_;
    size_t next = ((rank_ + 1) * NAF_) / NN_;
    offset_ = first;
    size_ = next - first;
}
size_t DivAssignment::operator[](size_t index) const {
    if (index < size_) return offset_ + index;
    throw Error("Assignment out of bounds, operator[]");
}
size_t DivAssignment::max() const {
    size_t m = NAF_ / NN_;
    if ((NAF_ % NN_) != 0) m += 1;
    return m;
}
bool DivAssignment::contains(size_t i) const { return i >= offset_ && i < offset_ + size_; }
size_t DivAssignment::index(size_t i) const {
    if (!contains(i)) throw Error("Assignment: index not contained");
    return i - offset_;
}

ModAssignment::ModAssignment(size_t NN, size_t rank, size_t NAF) : NN_(NN), rank_(rank), NAF_(NAF) {
    if (NN == 0) throw Error("ModAssignment: NN == 0");
    size_ = NAF_ / NN_;
    if ((NAF_ % NN_) != 0) {
        if (rank < (NAF_ - NN_ * (NAF_ / NN_))) size_ += 1;
    }
    offset_ = rank;
}
size_t ModAssignment::operator[](size_t index) const {
    if (index < size_) return offset_ + index * NN_;
    throw Error("Assignment out of bounds, operator[]");
}
size_t ModAssignment::max() const {
    size_t m = NAF_ / NN_;
    if ((NAF_ % NN_) != 0) m += 1;
    return m;
}
bool ModAssignment::contains(size_t i) const {
    if (i < offset_) return false;
    return ((i - offset_) % NN_) == 0;
}
size_t ModAssignment::index(size_t i) const {
    if (!contains(i)) throw Error("Assignment: index not contained");
    return (i - offset_) / NN_;
}

// decomposition_plan.cpp:29-66
DecompositionParameters::DecompositionParameters(size_t NN, size_t NQ, size_t NAF, size_t NNpP, size_t elbytesize,
                                                 bool replicated) {
    size_t NP = NN / NNpP;
    size_t NPused = NP;
    if (NQ < NP) NPused = NQ;
    size_t NNnotused = NN - NPused * NNpP;
    size_t NQcycles = ((NQ % NP) == 0) ? NQ / NP : NQ / NP + 1;
    size_t NAFcycles = ((NAF % NNpP) == 0) ? NAF / NNpP : NAF / NNpP + 1;
    size_t penalty = NNnotused * NQcycles * NAFcycles;
    penalty += (NPused * NQcycles - NQ) * (NNpP * NAFcycles);
    penalty += (NNpP * NAFcycles - NAF) * NQ;
    m_penalty = penalty;
    m_NAFcycles = NAFcycles;
    m_NQcycles = NQcycles;
    m_NNpP = NNpP;
    m_NN = NN;
    m_NQ = NQ;
    m_NAF = NAF;
    m_NP = NP;
    m_elbytesize = elbytesize;
    // replicated: the device stages ALL frames on every rank of a partition (multipole devices, the vector-sharded coherent
    // path), so a larger partition does not shrink a rank's share of the coordinates
    m_nbytesize = (replicated ? NAF : NAFcycles) * m_elbytesize;
}

// decomposition_plan.cpp:69-156
DecompositionPlan::DecompositionPlan(size_t nn, size_t nq, size_t naf, size_t elbytesize, size_t nmaxbytesize,
                                     const DecompositionLimits &lim, bool replicated) {
    if (naf < 1) throw Error("No data to decompose.");
    if (nn < 1) throw Error("No nodes to decompose onto.");
    if (lim.partitions_automatic) {
        size_t npmax = naf;
        if (naf > nn) npmax = nn;
        for (size_t nnpp = npmax; nnpp >= 1; nnpp--) {
            std::unique_ptr<DecompositionParameters> p_dp(new DecompositionParameters(nn, nq, naf, nnpp, elbytesize, replicated));
            if (p_dp->nbytesize() > nmaxbytesize) continue;
            if (!p_dp_best || p_dp->penalty() < p_dp_best->penalty()) p_dp_best = std::move(p_dp);
        }
        if (!p_dp_best)
            throw Error(
                "Automatic decomposition failed to match the necessary requirements. Either change the partition size "
                "manually or change the number of nodes. (limits.stage.memory.data)");
    } else {
        size_t nnpp = lim.partitions_size;
        if (nnpp > nn) nnpp = nn;
        if (lim.partitions_size > naf) nnpp = naf;
        if (nnpp < 1) nnpp = 1;
        p_dp_best.reset(new DecompositionParameters(nn, nq, naf, nnpp, elbytesize, replicated));
        // a manual partition size is not searched: say up front what the stager would say after reading the trajectory
        if (replicated && p_dp_best->nbytesize() > nmaxbytesize)
            throw Error("Insufficient Buffer size for coordinates (limits.memory.data): this device stages all frames on every "
                        "rank. Requested (bytes): " + std::to_string(p_dp_best->nbytesize()));
    }
    if (utilization() < lim.utilization) {
        p_dp_best.reset();
        throw Error("Utilization too low. Aborting. Change the number of nodes or the threshold value");
    }
}

double DecompositionPlan::utilization() const {
    size_t used = p_dp_best->get_NQ() * p_dp_best->get_NAF();
    size_t wasted = p_dp_best->penalty();
    return used * 1.0 / (used + wasted);
}

// ---------------------------------------------------------------------------------------------------------------
// generators (parameters.cpp:930-1189)
// ---------------------------------------------------------------------------------------------------------------
std::vector<CartesianCoor3D> create_from_scans(const std::vector<ScatteringVectorsScan> &scans) {
    if (scans.size() > 3) throw Error("More than 3 scan definitions are not supported.");
    std::vector<std::vector<CartesianCoor3D>> qvectors(scans.size());
    for (size_t i = 0; i < scans.size(); ++i) {
        const ScatteringVectorsScan &s = scans[i];
        if (s.points == 0) continue;
        if (s.points == 1) {
            double scal = (s.from + s.to) / 2;
            qvectors[i].push_back(scal * s.basevector);
            continue;
        }
        if (s.points == 2) {
            qvectors[i].push_back(s.from * s.basevector);
            qvectors[i].push_back(s.to * s.basevector);
            continue;
        }
        qvectors[i].push_back(s.from * s.basevector);
        for (size_t j = 1; j < (s.points - 1); j++) {
            // float-rounded fraction, as the reference's powf call (parameters.cpp:1151)
            double scal = s.from + powf((float)(j * 1.0 / (s.points - 1)), (float)s.exponent) * (s.to - s.from);
            qvectors[i].push_back(scal * s.basevector);
        }
        qvectors[i].push_back(s.to * s.basevector);
    }
    std::vector<CartesianCoor3D> out;
    if (scans.size() == 1) out = qvectors[0];
    if (scans.size() == 2)
        for (auto &a : qvectors[0])
            for (auto &b : qvectors[1]) out.push_back(a + b);
    if (scans.size() == 3)
        for (auto &a : qvectors[0])
            for (auto &b : qvectors[1])
                for (auto &c : qvectors[2]) out.push_back(a + b + c);
    return out;
}


void OrientationVectorsParameters::create() {
    if (type == "file") {
        for (auto &q : vectors) {
            double ql = q.length();
            if (ql != 0) q = (1.0 / ql) * q;
        }
    } else if (type == "sphere") {
        if (algorithm != "boost_uniform_on_sphere") throw Error("Vectors algorithm not understood: " + algorithm);
        vectors.clear();
        UniformOnSphere s(seed, 3);
        for (size_t i = 0; i < resolution; ++i) {
            std::vector<double> r = s();
            vectors.push_back(CartesianCoor3D(r[0], r[1], r[2]));
        }
    } else if (type == "cylinder") {
        vectors.clear();
        if (algorithm == "boost_uniform_on_sphere") {
            UniformOnSphere s(seed, 2);
            for (size_t i = 0; i < resolution; ++i) {
                std::vector<double> r = s();
                vectors.push_back(CartesianCoor3D(r[0], r[1], 0));
            }
        } else if (algorithm == "raster_linear") {
            const double M_2PI = 2 * M_PI;
            const double radincr = (M_2PI) / (360 * resolution);
            for (double phi = 0; phi < M_2PI; phi += radincr) vectors.push_back(CartesianCoor3D(cos(phi), sin(phi), 0));
        } else {
            throw Error("Vectors algorithm not understood: " + algorithm);
        }
    } else {
        throw Error("Vectors orientation averaging type not understood: " + type);
    }
}

void OrientationMultipoleParameters::create() {
    if (moments_type == "file") {
        // moments already given
    } else if (moments_type == "resolution") {
        moments.clear();
        if (type == "sphere") {
            moments.push_back(std::make_pair(0L, 0L));
            for (long l = 1; l <= resolution; ++l)
                for (long m = -l; m <= l; ++m) moments.push_back(std::make_pair(l, m));
        } else if (type == "cylinder") {
            moments.push_back(std::make_pair(0L, 0L));
            for (long l = 1; l <= resolution; ++l)
                for (long m = 0; m <= 3; ++m) moments.push_back(std::make_pair(l, m));
        } else {
            throw Error("Type not understood: scattering.average.orientation.multipole.moments.type=" + moments_type);
        }
    } else {
        throw Error("Type not understood: scattering.average.orientation.multipole.type=" + type);
    }
    // validity checks, parameters.cpp:1081-1118
    for (auto &mm : moments) {
        if (mm.first < 0) throw Error("Major multipole moment must be >= 0!");
        if (type == "cylinder") {
            if (mm.second < 0 || mm.second > 3) throw Error("Minor multipole moment must be between 0 and 3!");
            if (mm.first == 0 && mm.second != 0) throw Error("Minor multipole moment must be 0 for Major 0!");
        } else if (type == "sphere") {
            if (labs(mm.second) > mm.first) throw Error("Minor multipole moment must be between -Major and +Major!");
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// communicator / backend / timer
// ---------------------------------------------------------------------------------------------------------------
CallbackCommunicator::~CallbackCommunicator() {
    if (owned_ && v_.release) v_.release(v_.user);
}
void CallbackCommunicator::allreduce_sum(double *d, size_t n) {
    if (v_.allreduce_sum(v_.user, d, n) != 0) throw Error("communicator: allreduce failed");
}
void CallbackCommunicator::barrier() {
    if (v_.barrier(v_.user) != 0) throw Error("communicator: barrier failed");
}
std::shared_ptr<ICommunicator> CallbackCommunicator::split(int color) {
    void *u = v_.split(v_.user, color);
    if (!u) throw Error("communicator: split failed");
    sass_comm_vtbl nv = v_;
    nv.user = u;
    return std::make_shared<CallbackCommunicator>(nv, true);
}

const SgpuBackend &default_backend() {
    static const SgpuBackend be = {sgpu_init,
                                   sgpu_destroy,
                                   sgpu_last_error,
                                   sgpu_synchronize,
                                   sgpu_stage_frames,
                                   sgpu_frames_to_spherical,
                                   sgpu_stage_atoms,
                                   sgpu_stage_atoms_from_frames,
                                   sgpu_set_factors,
                                   sgpu_partial_len,
                                   sgpu_compute_all_vectors_partial,
                                   sgpu_compute_self_vectors_partial,
                                   sgpu_compute_mpsphere_partial,
                                   sgpu_finalize,
                                   sgpu_device_alloc,
                                   sgpu_device_free,
                                   sgpu_set_factors_batch,
                                   sgpu_mpsphere_amplitudes,
                                   sgpu_mpsphere_dsp_partial,
                                   sgpu_set_frame_window,
                                   sgpu_all_vectors_amplitudes,
                                   sgpu_all_vectors_dsp_partial,
                                   sgpu_compute_all_vectors_scan_partial,
                                   sgpu_all_vectors_scan_amplitudes,
                                   sgpu_stage_atoms_wave,
                                   sgpu_accumulate,
                                   sgpu_frames_to_cylindrical,
                                   sgpu_mpcylinder_amplitudes,
                                   sgpu_stage_atoms_prefetch,
                                   sgpu_stage_atoms_swap,
                                   sgpu_host_alloc,
                                   sgpu_host_free,
                                   sgpu_comm_adopt,
                                   sgpu_compute_all_vectors_scan_sharded,
                                   sgpu_compute_all_vectors_sharded};
    return be;
}

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
void Timer::start(const std::string &key) { start_[key] = now_s(); }
void Timer::stop(const std::string &key) {
    auto it = start_.find(key);
    if (it == start_.end()) return;
    auto &a = acc_[key];
    a.first += now_s() - it->second;
    a.second += 1;
}
double Timer::sum(const std::string &key) const {
    auto it = acc_.find(key);
    return it == acc_.end() ? 0.0 : it->second.first;
}
size_t Timer::count(const std::string &key) const {
    auto it = acc_.find(key);
    return it == acc_.end() ? 0 : it->second.second;
}
std::vector<std::string> Timer::keys() const {
    std::vector<std::string> k;
    for (auto &kv : acc_) k.push_back(kv.first);
    return k;
}

// ---------------------------------------------------------------------------------------------------------------
// stager
// ---------------------------------------------------------------------------------------------------------------
DataStagerByFrame::DataStagerByFrame(Sample &sample, ICommunicator &allcomm, ICommunicator &partitioncomm, Timer &timer,
                                     const SgpuBackend &be, sgpu_ctx *ctx, const Params &params)
    : m_sample(sample), allcomm_(allcomm), partitioncomm_(partitioncomm), timer_(timer), be_(be), ctx_(ctx), params_(params) {}

void DataStagerByFrame::stage_block() {
    DivAssignment mine(partitioncomm_.size(), partitioncomm_.rank(), m_sample.NF);
    size_t data_bytesize = mine.max() * m_sample.NA * 3 * sizeof(float);  // data_stager.cpp:57-63
    if (params_.limits.stage_memory_data < data_bytesize)
        throw Error("Insufficient Buffer size for coordinates (limits.memory.data) Requested (bytes): " +
                    std::to_string(data_bytesize));
    if (mine.size() == 0) throw Error("frame decomposition left a rank without frames");
    timer_.start("st:first");
    int rc = be_.stage_frames(ctx_, m_sample.frames + mine.offset() * m_sample.NA * 3, mine.size(), m_sample.NA,
                              SGPU_REPR_CARTESIAN);
    if (rc) throw Error(std::string("stage_frames: ") + be_.last_error(ctx_));
    rc = be_.set_frame_window(ctx_, m_sample.NF, mine.offset());
    if (rc) throw Error(std::string("set_frame_window: ") + be_.last_error(ctx_));
    timer_.stop("st:first");
    timer_.start("st:wait");
    allcomm_.barrier();
    timer_.stop("st:wait");
    if (params_.stager.dump) write(params_.stager.filepath, params_.stager.format);
}

void DataStagerByFrame::write(const std::string &filename, const std::string &format) {
    DivAssignment assignment(partitioncomm_.size(), partitioncomm_.rank(), m_sample.NF);
    if (allcomm_.rank() < partitioncomm_.size()) {  // the first partition writes
        timer_.start("st:dump");
        if (format != "dcd") throw Error("Format for coordinate dumping not known: " + format);
        DCDCoordinateWriter cw(filename, m_sample.NF, m_sample.NA);
        if (allcomm_.rank() == 0) cw.init();
        partitioncomm_.barrier();
        cw.prepare();
        partitioncomm_.barrier();
        for (size_t c = 0; c < assignment.max(); ++c) {  // consecutive blocks, one frame per rank and round
            if (c < assignment.size()) cw.write(m_sample.frames + assignment[c] * m_sample.NA * 3, assignment[c], 1);
            partitioncomm_.barrier();
        }
        timer_.stop("st:dump");
    }
    timer_.start("st:wait");
    allcomm_.barrier();
    timer_.stop("st:wait");
}

void DataStagerByFrame::stage(int repr) {
    // every GPU of the partition holds all frames
    size_t data_bytesize = m_sample.NF * m_sample.NA * 3 * sizeof(float);
    if (params_.limits.stage_memory_data < data_bytesize)
        throw Error("Insufficient Buffer size for coordinates (limits.memory.data) Requested (bytes): " +
                    std::to_string(data_bytesize));
    timer_.start("st:first");
    int rc = be_.stage_frames(ctx_, m_sample.frames, m_sample.NF, m_sample.NA, SGPU_REPR_CARTESIAN);
    if (rc) throw Error(std::string("stage_frames: ") + be_.last_error(ctx_));
    if (repr == SGPU_REPR_SPHERICAL) {
        rc = be_.frames_to_spherical(ctx_);
        if (rc) throw Error(std::string("frames_to_spherical: ") + be_.last_error(ctx_));
    }
    timer_.stop("st:first");
    timer_.start("st:wait");
    allcomm_.barrier();
    timer_.stop("st:wait");
    if (params_.stager.dump) write(params_.stager.filepath, params_.stager.format);  // the cartesian frames, as the reference's
}

DataStagerByAtom::DataStagerByAtom(Sample &sample, ICommunicator &allcomm, ICommunicator &partitioncomm, Timer &timer,
                                   const SgpuBackend &be, sgpu_ctx *ctx, const Params &params)
    : m_sample(sample), allcomm_(allcomm), partitioncomm_(partitioncomm), timer_(timer), be_(be), ctx_(ctx), params_(params) {
    ModAssignment assignment(partitioncomm_.size(), partitioncomm_.rank(), m_sample.NA);
    size_t data_bytesize = assignment.max() * m_sample.NF * 3 * sizeof(float);  // data_stager.cpp:194-204
    if (params_.limits.stage_memory_data < data_bytesize)
        throw Error("Insufficient Buffer size for coordinates (limits.memory.data) Requested (bytes): " +
                    std::to_string(data_bytesize));
}

void DataStagerByAtom::stage() {
    timer_.start("st:first");
    int rc = be_.stage_atoms_from_frames(ctx_, m_sample.frames, m_sample.NF, m_sample.NA, partitioncomm_.size(),
                                         partitioncomm_.rank());
    if (rc) throw Error(std::string("stage_atoms_from_frames: ") + be_.last_error(ctx_));
    timer_.stop("st:first");
    timer_.start("st:wait");
    allcomm_.barrier();
    timer_.stop("st:wait");
    if (params_.stager.dump) write(params_.stager.filepath, params_.stager.format);
}

void DataStagerByAtom::write(const std::string &filename, const std::string &format) {
    ModAssignment assignment(partitioncomm_.size(), partitioncomm_.rank(), m_sample.NA);
    if (allcomm_.rank() < partitioncomm_.size()) {
        timer_.start("st:dump");
        if (format != "dcd") throw Error("Format for coordinate dumping not known: " + format);
        DCDCoordinateWriter cw(filename, m_sample.NA, m_sample.NF);  // blocks = atoms, entries = frames
        if (partitioncomm_.rank() == 0) cw.init();
        partitioncomm_.barrier();
        cw.prepare();
        partitioncomm_.barrier();
        std::vector<float> line(m_sample.NF * 3);
        for (size_t c = 0; c < assignment.max(); ++c) {
            if (c < assignment.size()) {
                const size_t atom = assignment[c];
                for (size_t f = 0; f < m_sample.NF; f++)
                    for (int k = 0; k < 3; k++) line[3 * f + k] = m_sample.frames[(f * m_sample.NA + atom) * 3 + k];
                cw.write(line.data(), atom, 1);
            }
            partitioncomm_.barrier();
        }
        timer_.stop("st:dump");
    }
    timer_.start("st:wait");
    allcomm_.barrier();
    timer_.stop("st:wait");
}

// ---------------------------------------------------------------------------------------------------------------
// AbstractScatterDevice (abstract_scatter_device.cpp)
// ---------------------------------------------------------------------------------------------------------------
AbstractScatterDevice::AbstractScatterDevice(std::shared_ptr<ICommunicator> allcomm,
                                             std::shared_ptr<ICommunicator> partitioncomm, Sample &sample,
                                             std::vector<CartesianCoor3D> vectors, size_t NAF, IResultSink *sink,
                                             const Params &params, const SgpuBackend &be, sgpu_ctx *ctx)
    : allcomm_(allcomm), partitioncomm_(partitioncomm), sample_(sample), vectors_(vectors), p_hdf5writer_(sink),
      params_(params), be_(be), ctx_(ctx), afinal_(0), a2final_(0) {
    (void)NAF;
    NN = allcomm_->size();
    NA = sample_.NA;
    NF = sample_.NF;
    if (!ctx_) {
        int rc = be_.init(-1, &ctx_);  // -1: "the device this process is bound to" (current CUDA device)
        if (rc) throw Error(std::string("sgpu_init: ") + be_.last_error(nullptr));
        own_ctx_ = true;
    }
}

AbstractScatterDevice::~AbstractScatterDevice() {
    if (d_partial_) be_.device_free(d_partial_);
    if (own_ctx_ && ctx_) be_.destroy(ctx_);
}

void AbstractScatterDevice::ck(int rc, const char *what) {
    if (rc) throw Error(std::string(what) + ": " + be_.last_error(ctx_));
}

int AbstractScatterDevice::dsp_type_code() const {
    const std::string &t = params_.scattering.dsp_type;
    if (t == "autocorrelate") return SGPU_DSP_AUTOCORRELATE;
    if (t == "square") return SGPU_DSP_SQUARE;
    if (t == "plain") return SGPU_DSP_PLAIN;
    // all_vectors_scatter_device.cpp:224-228
    throw Error("DSP type not understood: " + t + " (scattering.dsp.type == autocorrelate, square, plain)");
}

int AbstractScatterDevice::dsp_method_code() const {
    const std::string &m = params_.scattering.dsp_method;
    if (m == "fftw") return SGPU_METHOD_FFTW;
    if (m == "direct") return SGPU_METHOD_DIRECT;
    if (params_.scattering.dsp_type != "autocorrelate") return SGPU_METHOD_FFTW;
    throw Error("Correlation method not understood (scattering.dsp.method == direct, fftw)");  // :218-221
}

double *AbstractScatterDevice::partial_buffer(int dsp_type) {
    size_t len = 0;
    ck(be_.partial_len(ctx_, dsp_type, &len), "sgpu_partial_len");
    if (len > partial_cap_) {
        if (d_partial_) be_.device_free(d_partial_);
        d_partial_ = nullptr;
        void *p = nullptr;
        if (be_.device_alloc(&p, len * sizeof(double))) throw Error("device allocation of the partial buffer failed");
        d_partial_ = static_cast<double *>(p);
        partial_cap_ = len;
    }
    return d_partial_;
}

// the three boost::mpi::reduce calls of compute() (all_vectors_scatter_device.cpp:324-352) as one all-reduce of the
// packed partial, followed by the inverse transform / scaling on every rank (rank 0 of the partition writes)
void AbstractScatterDevice::reduce_and_finalize(int dsp_type, double scale, double *d_partial) {
    if (!d_partial) d_partial = d_partial_;
    size_t len = 0;
    ck(be_.partial_len(ctx_, dsp_type, &len), "sgpu_partial_len");
    timer_.start("sd:c:wait");
    ck(be_.synchronize(ctx_), "sgpu_synchronize");
    timer_.stop("sd:c:wait");
    timer_.start("sd:c:reduce");
    if (partitioncomm_->size() > 1) partitioncomm_->allreduce_sum(d_partial, len);
    timer_.stop("sd:c:reduce");
    double af[2], a2f[2];
    ck(be_.finalize(ctx_, d_partial, dsp_type, dsp_method_code(), scale, atfinal_.data(), af, a2f), "sgpu_finalize");
    afinal_ = std::complex<double>(af[0], af[1]);
    a2final_ = std::complex<double>(a2f[0], a2f[1]);
}

bool AbstractScatterDevice::ram_check() {
    // the reference checks limits.computation.memory.* against host buffers; here the device allocator reports
    // exhaustion (SGPU_ENOMEM) at stage/compute time, and the stagers check limits.stage.memory.data
    return true;
}

void AbstractScatterDevice::run() {
    if (!ram_check()) throw terminate_request();
    print_pre_stage_info();
    allcomm_->barrier();
    timer_.start("sd:stage");
    stage_data();
    timer_.stop("sd:stage");
    allcomm_->barrier();
    print_post_stage_info();

    atfinal_.assign(2 * NF, 0.0);

    print_pre_runner_info();
    allcomm_->barrier();
    timer_.start("sd:runner");
    runner();
    timer_.stop("sd:runner");
    allcomm_->barrier();
    print_post_runner_info();
}

void AbstractScatterDevice::runner() {
    while (status() == 0) {
        timer_.start("sd:compute");
        compute();
        timer_.stop("sd:compute");
        timer_.start("sd:write");
        write();
        timer_.stop("sd:write");
        next();
    }
}

void AbstractScatterDevice::next() {
    if (current_vector_ >= vectors_.size()) return;
    current_vector_++;
}

double AbstractScatterDevice::progress() {
    double scale = 1.0 / vectors_.size();
    return current_vector_ * scale;
}

size_t AbstractScatterDevice::status() { return current_vector_ == vectors_.size() ? 1 : 0; }

void AbstractScatterDevice::write() {
    if (partitioncomm_->rank() == 0 && p_hdf5writer_) {
        CartesianCoor3D vector = vectors_[current_vector_];
        p_hdf5writer_->write(vector, atfinal_.data(), NF, afinal_, a2final_);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// AbstractVectorsScatterDevice (abstract_vectors_scatter_device.cpp:96-175)
// ---------------------------------------------------------------------------------------------------------------
double AbstractVectorsScatterDevice::progress() {
    double scale1 = 1.0 / vectors_.size();
    double scale2 = 1.0 / NM;
    return current_vector_ * scale1 + current_subvector_ * scale1 * scale2;
}

void AbstractVectorsScatterDevice::init_subvectors(CartesianCoor3D &q) {
    subvector_index_.clear();
    const OrientationVectorsParameters &ov = params_.scattering.vectors;
    // the reference keys on vectors.size()>0, which is only the case for orientation.type=="vectors"
    if (params_.scattering.orientation_type == "vectors" && ov.vectors.size() > 0) {
        if (ov.type == "file" || ov.type == "sphere") {
            double ql = q.length();
            for (size_t i = 0; i < ov.vectors.size(); ++i) subvector_index_.push_back(ql * ov.vectors[i]);
        } else if (ov.type == "cylinder") {
            CartesianCoor3D o = params_.scattering.axis;
            CartesianVectorBase base(o);
            CartesianCoor3D qprojected = base.project(q);
            // CylinderCoor3D(CartesianCoor3D): r = sqrt(x^2+y^2), z = z (coor3d.cpp:116-140)
            double qr = std::sqrt(std::pow(qprojected.x, 2) + std::pow(qprojected.y, 2));
            double qz = qprojected.z;
            if (qr == 0) {
                subvector_index_.push_back(qprojected);  // quirk kept: the projected vector (:134-135)
            } else {
                for (size_t i = 0; i < ov.vectors.size(); ++i) {
                    const CartesianCoor3D &vec = ov.vectors[i];
                    CartesianCoor3D qnew = qz * base[2] + qr * (vec.x * base[0] + vec.y * base[1]);
                    subvector_index_.push_back(qnew);
                }
            }
        }
    } else {
        subvector_index_.push_back(q);
    }
    NM = subvector_index_.size();
}

// ---------------------------------------------------------------------------------------------------------------
// AllVectorsScatterDevice (all_vectors_scatter_device.cpp)
// ---------------------------------------------------------------------------------------------------------------
AllVectorsScatterDevice::~AllVectorsScatterDevice() {
    if (d_amp_) be_.device_free(d_amp_);
}

void AllVectorsScatterDevice::stage_data() {
    DataStagerByFrame data_stager(sample_, *allcomm_, *partitioncomm_, timer_, be_, ctx_, params_);
    const size_t NNPP = partitioncomm_->size();
    const std::string &mode = params_.limits.coherent_sharding;
    if (mode != "auto" && mode != "frames" && mode != "vectors")
        throw Error("limits.decomposition.coherent not understood: " + mode + " (auto, frames, vectors)");
    // frames: per |q| the ranks exchange A[NM][NF] (16 B per entry) against NA*NF*NM/NNPP evaluations each, so the
    // exchange is amortised once the sample has some hundred atoms per rank; below that keep the frames replicated
    frame_sharded_ = NNPP > 1 && NF >= NNPP && (mode == "frames" || (mode == "auto" && NA >= 64 * NNPP));
    if (frame_sharded_) data_stager.stage_block();
    else data_stager.stage(SGPU_REPR_CARTESIAN);
    factors_.assign(NA, 0.0);
    // a partition communicator that is an NCCL communicator goes to the library: the amplitude exchange then runs as
    // grouped ncclSend/ncclRecv on the device streams (all_to_all + alignpad, all_vectors_scatter_device.cpp:169-207)
    nccl_sharded_ = false;
    if (frame_sharded_ && partitioncomm_->nccl_comm() && be_.comm_adopt && be_.compute_all_vectors_scan_sharded &&
        be_.compute_all_vectors_sharded) {
        ck(be_.comm_adopt(ctx_, partitioncomm_->nccl_comm(), (int)NNPP, (int)partitioncomm_->rank()), "sgpu_comm_adopt");
        nccl_sharded_ = true;
    }
}

// Frame decomposition (the reference's own, all_vectors_scatter_device.cpp:245-361): amplitudes of this rank's frames
// for ALL subvectors, exchange (there: all_to_all per block of NNPP subvectors, :170-205; here: one sum of the
// zero-padded A over NVSwitch), then every rank correlates DivAssignment(NNPP, rank, NM) of the timelines.
void AllVectorsScatterDevice::compute_frame_sharded() {
    const int dsp = dsp_type_code();
    dsp_method_code();
    std::vector<double> qv(3 * NM);
    for (size_t i = 0; i < NM; i++) {
        const CartesianCoor3D &s = subvector_index_[i];
        qv[3 * i] = s.x;
        qv[3 * i + 1] = s.y;
        qv[3 * i + 2] = s.z;
    }
    if (nccl_sharded_) {
        // amplitudes, exchange, DSP of this rank's timelines and the reduction of the packed partial inside the library
        double *partial = partial_buffer(dsp);
        timer_.start("sd:c:block");
        ck(be_.compute_all_vectors_sharded(ctx_, qv.data(), NM, dsp, partial), "sgpu_compute_all_vectors_sharded");
        timer_.stop("sd:c:block");
        timer_.start("sd:c:b:exchange");
        ck(be_.synchronize(ctx_), "sgpu_synchronize");
        timer_.stop("sd:c:b:exchange");
        current_subvector_ = NM;
        double af[2], a2f[2];
        ck(be_.finalize(ctx_, partial, dsp, dsp_method_code(), 1.0 / subvector_index_.size(), atfinal_.data(), af, a2f), "sgpu_finalize");
        afinal_ = std::complex<double>(af[0], af[1]);
        a2final_ = std::complex<double>(a2f[0], a2f[1]);
        return;
    }
    const size_t amp_len = 2 * NM * NF;
    if (amp_len > amp_cap_) {
        if (d_amp_) be_.device_free(d_amp_);
        d_amp_ = nullptr;
        void *pp = nullptr;
        if (be_.device_alloc(&pp, amp_len * sizeof(double))) throw Error("device allocation of the amplitude buffer failed");
        d_amp_ = static_cast<double *>(pp);
        amp_cap_ = amp_len;
    }
    timer_.start("sd:c:block");
    ck(be_.all_vectors_amplitudes(ctx_, qv.data(), NM, d_amp_), "sgpu_all_vectors_amplitudes");
    timer_.stop("sd:c:block");
    timer_.start("sd:c:wait");
    ck(be_.synchronize(ctx_), "sgpu_synchronize");
    timer_.stop("sd:c:wait");
    timer_.start("sd:c:b:exchange");
    partitioncomm_->allreduce_sum(d_amp_, amp_len);
    timer_.stop("sd:c:b:exchange");
    DivAssignment mine(partitioncomm_->size(), partitioncomm_->rank(), NM);
    double *partial = partial_buffer(dsp);
    timer_.start("sd:c:b:dspstore");
    ck(be_.all_vectors_dsp_partial(ctx_, d_amp_, mine.offset(), mine.size(), dsp, partial), "sgpu_all_vectors_dsp_partial");
    timer_.stop("sd:c:b:dspstore");
    current_subvector_ = NM;
    reduce_and_finalize(dsp, 1.0 / subvector_index_.size());
}

void AllVectorsScatterDevice::compute() {
    CartesianCoor3D q = vectors_[current_vector_];
    timer_.start("sd:c:init");
    init_subvectors(q);
    sample_.factors(q.length(), factors_.data());  // scatterfactors.update(q) (:245)
    ck(be_.set_factors(ctx_, factors_.data(), NA), "sgpu_set_factors");
    timer_.stop("sd:c:init");

    current_subvector_ = 0;
    if (frame_sharded_) {
        compute_frame_sharded();
        return;
    }
    const int dsp = dsp_type_code();
    dsp_method_code();
    // q-vector decomposition inside the partition: rank r takes DivAssignment(NNPP, r, NM) of the subvectors
    // (every GPU holds all frames; no amplitude exchange, only the packed partial is summed)
    DivAssignment mine(partitioncomm_->size(), partitioncomm_->rank(), NM);
    std::vector<double> qv(3 * std::max<size_t>(mine.size(), 1));
    for (size_t i = 0; i < mine.size(); i++) {
        const CartesianCoor3D &s = subvector_index_[mine[i]];
        qv[3 * i] = s.x;
        qv[3 * i + 1] = s.y;
        qv[3 * i + 2] = s.z;
    }
    double *partial = partial_buffer(dsp);
    timer_.start("sd:c:block");
    ck(be_.compute_all_vectors_partial(ctx_, qv.data(), mine.size(), dsp, partial), "sgpu_compute_all_vectors_partial");
    current_subvector_ = NM;
    timer_.stop("sd:c:block");
    reduce_and_finalize(dsp, 1.0 / subvector_index_.size());  // factor = 1/NM (:355-360)
}

// Longest run of q-vectors starting at `first` that shares its directions: subvectors(q_n) = |q_n| v_m for all n, m.
// Verified numerically on the subvectors the reference's init_subvectors produces, so every orientation type
// (sphere, file, cylinder, none) qualifies exactly when it has that structure; the spacing of the |q_n| is the
// backend's business.  Returns 1 when there is no such run.
size_t AllVectorsScatterDevice::scan_length(size_t first, std::vector<double> &v, std::vector<double> &s) {
    const size_t maxn = std::min<size_t>(params_.limits.coherent_scan, vectors_.size() - first);
    if (maxn < 4) return 1;  // a pass pays for two sincos per direction: not worth it below 4 |q|
    CartesianCoor3D q0 = vectors_[first];
    const double s0 = q0.length();
    if (s0 <= 0.0) return 1;
    init_subvectors(q0);
    const size_t nm = NM;
    if (nm < 8) return 1;  // a CTA serves 8-12 directions; below that the per-|q| kernel is the better fit
    v.resize(3 * nm);
    for (size_t m = 0; m < nm; m++) {
        v[3 * m] = subvector_index_[m].x / s0;
        v[3 * m + 1] = subvector_index_[m].y / s0;
        v[3 * m + 2] = subvector_index_[m].z / s0;
    }
    s.assign(1, s0);
    // device memory for the amplitudes of the batch: at most ~1/8 of the coordinate budget
    const size_t per_q = nm * NF * 2 * sizeof(double);
    const size_t cap = std::max<size_t>(1, (params_.limits.stage_memory_data / 8) / std::max<size_t>(per_q, 1));
    for (size_t n = 1; n < maxn && n < cap; n++) {
        CartesianCoor3D qn = vectors_[first + n];
        const double sn = qn.length();
        if (sn <= 0.0) break;
        init_subvectors(qn);
        if (NM != nm) break;
        const double tol = 4e-15 * sn;
        bool ok = true;
        for (size_t m = 0; m < nm && ok; m++) {
            const CartesianCoor3D &sv = subvector_index_[m];
            ok = std::fabs(sv.x - sn * v[3 * m]) <= tol && std::fabs(sv.y - sn * v[3 * m + 1]) <= tol &&
                 std::fabs(sv.z - sn * v[3 * m + 2]) <= tol;
        }
        if (!ok) break;
        s.push_back(sn);
    }
    if (s.size() < 4) return 1;
    if (params_.limits.coherent_scan_snap) {
        const size_t n = s.size();
        const double ds = (s[n - 1] - s[0]) / (double)(n - 1);
        double dev = 0.0, smax = 0.0;
        for (size_t i = 0; i < n; i++) {
            dev = std::max(dev, std::fabs(s[i] - (s[0] + (double)i * ds)));
            smax = std::max(smax, s[i]);
        }
        if (dev > 0.0 && dev <= 1e-6 * smax) {
            for (size_t i = 0; i < n; i++) {
                const double snew = s[0] + (double)i * ds;
                vectors_[first + i] = (snew / s[i]) * vectors_[first + i];  // the snapped vector is what gets written
                s[i] = snew;
            }
        }
    }
    return s.size();
}

void AllVectorsScatterDevice::compute_scan(size_t nq, const std::vector<double> &v, const std::vector<double> &s) {
    const size_t nm = v.size() / 3;
    NM = nm;
    timer_.start("sd:c:init");
    std::vector<double> fb(nq * NA);
    for (size_t n = 0; n < nq; n++) {
        CartesianCoor3D q = vectors_[current_vector_ + n];
        sample_.factors(q.length(), &fb[n * NA]);  // scatterfactors.update(q) per |q|
    }
    ck(be_.set_factors_batch(ctx_, fb.data(), nq, NA), "sgpu_set_factors_batch");
    timer_.stop("sd:c:init");
    const int dsp = dsp_type_code();
    dsp_method_code();
    size_t plen = 0;
    ck(be_.partial_len(ctx_, dsp, &plen), "sgpu_partial_len");
    if (nq * plen > partial_cap_) {
        if (d_partial_) be_.device_free(d_partial_);
        d_partial_ = nullptr;
        void *pp = nullptr;
        if (be_.device_alloc(&pp, nq * plen * sizeof(double))) throw Error("device allocation of the partial buffer failed");
        d_partial_ = static_cast<double *>(pp);
        partial_cap_ = nq * plen;
    }
    const size_t NNPP = partitioncomm_->size();
    DivAssignment mine(NNPP, partitioncomm_->rank(), nm);
    bool reduced = false;
    if (frame_sharded_ && nccl_sharded_) {
        timer_.start("sd:c:block");
        ck(be_.compute_all_vectors_scan_sharded(ctx_, v.data(), nm, s.data(), nq, dsp, d_partial_),
           "sgpu_compute_all_vectors_scan_sharded");
        timer_.stop("sd:c:block");
        timer_.start("sd:c:b:exchange");
        ck(be_.synchronize(ctx_), "sgpu_synchronize");
        timer_.stop("sd:c:b:exchange");
        reduced = true;  // the library has summed the packed partials over the partition
    } else if (frame_sharded_) {
        const size_t amp_len = 2 * nq * nm * NF;
        if (amp_len > amp_cap_) {
            if (d_amp_) be_.device_free(d_amp_);
            d_amp_ = nullptr;
            void *pp = nullptr;
            if (be_.device_alloc(&pp, amp_len * sizeof(double))) throw Error("device allocation of the amplitude buffer failed");
            d_amp_ = static_cast<double *>(pp);
            amp_cap_ = amp_len;
        }
        timer_.start("sd:c:block");
        ck(be_.all_vectors_scan_amplitudes(ctx_, v.data(), nm, s.data(), nq, d_amp_), "sgpu_all_vectors_scan_amplitudes");
        timer_.stop("sd:c:block");
        timer_.start("sd:c:wait");
        ck(be_.synchronize(ctx_), "sgpu_synchronize");
        timer_.stop("sd:c:wait");
        timer_.start("sd:c:b:exchange");
        partitioncomm_->allreduce_sum(d_amp_, amp_len);
        timer_.stop("sd:c:b:exchange");
        timer_.start("sd:c:b:dspstore");
        for (size_t n = 0; n < nq; n++)
            ck(be_.all_vectors_dsp_partial(ctx_, d_amp_ + n * 2 * nm * NF, mine.offset(), mine.size(), dsp, d_partial_ + n * plen),
               "sgpu_all_vectors_dsp_partial");
        timer_.stop("sd:c:b:dspstore");
    } else {
        timer_.start("sd:c:block");
        ck(be_.compute_all_vectors_scan_partial(ctx_, v.data() + 3 * mine.offset(), mine.size(), s.data(), nq, dsp, d_partial_),
           "sgpu_compute_all_vectors_scan_partial");
        timer_.stop("sd:c:block");
    }
    timer_.start("sd:c:wait");
    ck(be_.synchronize(ctx_), "sgpu_synchronize");
    timer_.stop("sd:c:wait");
    timer_.start("sd:c:reduce");
    if (NNPP > 1 && !reduced) partitioncomm_->allreduce_sum(d_partial_, nq * plen);
    timer_.stop("sd:c:reduce");
    batch_atfinal_.assign(nq, std::vector<double>(2 * NF));
    batch_afinal_.assign(nq, 0.0);
    batch_a2final_.assign(nq, 0.0);
    for (size_t n = 0; n < nq; n++) {
        double af[2], a2f[2];
        ck(be_.finalize(ctx_, d_partial_ + n * plen, dsp, dsp_method_code(), 1.0 / (double)nm, batch_atfinal_[n].data(), af, a2f),
           "sgpu_finalize");
        batch_afinal_[n] = std::complex<double>(af[0], af[1]);
        batch_a2final_[n] = std::complex<double>(a2f[0], a2f[1]);
    }
    scans_++;
}

void AllVectorsScatterDevice::runner() {
    std::vector<double> v, s;
    while (status() == 0) {
        const size_t nq = scan_length(current_vector_, v, s);
        if (nq < 2) {
            timer_.start("sd:compute");
            compute();
            timer_.stop("sd:compute");
            timer_.start("sd:write");
            write();
            timer_.stop("sd:write");
            next();
            continue;
        }
        timer_.start("sd:compute");
        timer_.start("sd:c:scan");
        compute_scan(nq, v, s);
        timer_.stop("sd:c:scan");
        timer_.stop("sd:compute");
        for (size_t n = 0; n < nq; n++) {
            atfinal_ = batch_atfinal_[n];
            afinal_ = batch_afinal_[n];
            a2final_ = batch_a2final_[n];
            timer_.start("sd:write");
            write();
            timer_.stop("sd:write");
            next();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// SelfVectorsScatterDevice (self_vectors_scatter_device.cpp)
// ---------------------------------------------------------------------------------------------------------------
SelfVectorsScatterDevice::SelfVectorsScatterDevice(std::shared_ptr<ICommunicator> allcomm,
                                                   std::shared_ptr<ICommunicator> partitioncomm, Sample &sample,
                                                   std::vector<CartesianCoor3D> vectors, size_t NAF, IResultSink *sink,
                                                   const Params &params, const SgpuBackend &be, sgpu_ctx *ctx)
    : AbstractVectorsScatterDevice(allcomm, partitioncomm, sample, vectors, NAF, sink, params, be, ctx),
      assignment_(partitioncomm->size(), partitioncomm->rank(), NAF) {}

SelfVectorsScatterDevice::~SelfVectorsScatterDevice() {
    if (d_acc_) be_.device_free(d_acc_);
    if (h_atoms_) be_.host_free(h_atoms_);
}

// DataStagerByAtom for a share that does not fit the coordinate budget (data_stager.cpp:214-349 stops there, :194-204): this
// rank's atoms are gathered ONCE from the frame-major trajectory into an atom-major pinned host buffer [n_local][NF][3]
// (tiles of frames x atoms so that both sides move whole cache lines; host threads), from which contiguous blocks of atoms
// then stream to the GPU.
void SelfVectorsScatterDevice::gather_atoms_to_host() {
    const size_t nloc = assignment_.size();
    if (nloc == 0) return;
    void *p = nullptr;
    if (be_.host_alloc(&p, nloc * NF * 3 * sizeof(float))) throw Error("pinned host allocation of the rank's atoms failed");
    h_atoms_ = static_cast<float *>(p);
    const float *frames = sample_.frames;
    const size_t TF = 64, TA = 256;
    const size_t ntiles = (nloc + TA - 1) / TA;
    const unsigned nthreads = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), (unsigned)std::min<size_t>(ntiles, 16)));
    auto work = [&](unsigned tid) {
        for (size_t ta = tid; ta < ntiles; ta += nthreads) {
            const size_t a0 = ta * TA, a1 = std::min(nloc, a0 + TA);
            for (size_t f0 = 0; f0 < NF; f0 += TF) {
                const size_t f1 = std::min(NF, f0 + TF);
                for (size_t a = a0; a < a1; a++) {
                    const size_t src_atom = assignment_[a];
                    float *dst = h_atoms_ + (a * NF + f0) * 3;
                    for (size_t f = f0; f < f1; f++) {
                        const float *src = frames + (f * NA + src_atom) * 3;
                        dst[0] = src[0];
                        dst[1] = src[1];
                        dst[2] = src[2];
                        dst += 3;
                    }
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nthreads; t++) pool.emplace_back(work, t);
    work(0);
    for (auto &t : pool) t.join();
}

void SelfVectorsScatterDevice::stage_data() {
    factors_.assign(NA, 0.0);
    const size_t atom_bytes = NF * 3 * sizeof(float);
    const size_t share = assignment_.max() * atom_bytes;  // data_stager.cpp:194-204
    if (share > params_.limits.stage_memory_data && params_.limits.stage_stream) {
        // the share does not fit the coordinate budget: waves of as many atoms as do; staged inside runner()
        // two wave buffers live on the device: half the budget each
        streamed_ = true;
        wave_atoms_ = std::max<size_t>(1, params_.limits.stage_memory_data / (2 * atom_bytes));
        waves_ = (assignment_.size() + wave_atoms_ - 1) / wave_atoms_;
        timer_.start("sd:stage");
        gather_atoms_to_host();
        timer_.stop("sd:stage");
        return;
    }
    DataStagerByAtom data_stager(sample_, *allcomm_, *partitioncomm_, timer_, be_, ctx_, params_);
    data_stager.stage();
}

void SelfVectorsScatterDevice::compute_partial(size_t first, size_t count, int dsp, double *d_out) {
    CartesianCoor3D q = vectors_[current_vector_];
    timer_.start("sd:c:init");
    init_subvectors(q);
    sample_.factors(q.length(), factors_.data());
    // scatterfactors.get(assignment_[ai]) (:291): factors of the staged atoms in staged order
    std::vector<double> mine(std::max<size_t>(count, 1));
    for (size_t i = 0; i < count; i++) mine[i] = factors_[assignment_[first + i]];
    std::vector<double> qv(3 * NM);
    for (size_t i = 0; i < NM; i++) {
        qv[3 * i] = subvector_index_[i].x;
        qv[3 * i + 1] = subvector_index_[i].y;
        qv[3 * i + 2] = subvector_index_[i].z;
    }
    timer_.stop("sd:c:init");
    timer_.start("sd:c:block");
    if (count > 0) {
        ck(be_.set_factors(ctx_, mine.data(), count), "sgpu_set_factors");
        ck(be_.compute_self_vectors_partial(ctx_, qv.data(), NM, dsp, d_out), "sgpu_compute_self_vectors_partial");
    } else {
        // a rank without atoms contributes zeros: an empty q-list zeroes the partial
        ck(be_.compute_self_vectors_partial(ctx_, qv.data(), 0, dsp, d_out), "sgpu_compute_self_vectors_partial");
    }
    current_subvector_ = NM;
    timer_.stop("sd:c:block");
}

// Streamed runner: the reference's loop is |q|-outer with all of the rank's atoms resident (runner(),
// abstract_scatter_device.cpp:162-173).  When they are not, the loops are exchanged -- every wave of atoms is staged once
// and evaluated for all |q| of the partition -- and the packed partials, which are sums over atoms, accumulate per |q|.
// The all-reduce over the partition and the write happen per |q| after the last wave, in the reference's order.
void SelfVectorsScatterDevice::runner() {
    if (!streamed_) {
        AbstractScatterDevice::runner();
        return;
    }
    const int dsp = dsp_type_code();
    dsp_method_code();
    const size_t nq = vectors_.size();
    size_t plen = 0;
    auto wave_count = [&](size_t w) { return waves_ ? std::min(wave_atoms_, assignment_.size() - w * wave_atoms_) : (size_t)0; };
    auto prefetch = [&](size_t w) {
        if (w < waves_ && wave_count(w) > 0)
            ck(be_.stage_atoms_prefetch(ctx_, h_atoms_ + w * wave_atoms_ * NF * 3, wave_count(w), NF), "sgpu_stage_atoms_prefetch");
    };
    prefetch(0);
    for (size_t w = 0; w < std::max<size_t>(waves_, 1); w++) {
        const size_t first = w * wave_atoms_;
        const size_t count = wave_count(w);
        timer_.start("sd:stage");
        if (count > 0) {
            // wave w becomes the staged atoms; wave w+1 starts travelling while w is evaluated (double buffering)
            ck(be_.stage_atoms_swap(ctx_), "sgpu_stage_atoms_swap");
            prefetch(w + 1);
        }
        timer_.stop("sd:stage");
        if (w == 0) {
            partial_buffer(dsp);
            ck(be_.partial_len(ctx_, dsp, &plen), "sgpu_partial_len");
            void *p = nullptr;
            if (be_.device_alloc(&p, nq * plen * sizeof(double))) throw Error("device allocation of the wave accumulators failed");
            d_acc_ = static_cast<double *>(p);
        }
        timer_.start("sd:compute");
        for (current_vector_ = 0; current_vector_ < nq; current_vector_++) {
            double *acc = d_acc_ + current_vector_ * plen;
            // the first wave writes the accumulator, later waves add to it
            compute_partial(first, count, dsp, w == 0 ? acc : d_partial_);
            if (w > 0) ck(be_.accumulate(ctx_, acc, d_partial_, plen), "sgpu_accumulate");
        }
        timer_.stop("sd:compute");
    }
    for (current_vector_ = 0; current_vector_ < nq;) {
        init_subvectors(vectors_[current_vector_]);
        reduce_and_finalize(dsp, 1.0 / subvector_index_.size(), d_acc_ + current_vector_ * plen);  // :233-238
        timer_.start("sd:write");
        write();
        timer_.stop("sd:write");
        next();
    }
}

void SelfVectorsScatterDevice::compute() {
    const int dsp = dsp_type_code();
    dsp_method_code();
    compute_partial(0, assignment_.size(), dsp, partial_buffer(dsp));
    reduce_and_finalize(dsp, 1.0 / subvector_index_.size());  // :233-238
}

// ---------------------------------------------------------------------------------------------------------------
// MPSphereScatterDevice (multipole_scatter_device.cpp:32-498)
// ---------------------------------------------------------------------------------------------------------------
void MPSphereScatterDevice::init_moments(CartesianCoor3D &q) {
    multipole_index_.clear();
    for (size_t i = 0; i < params_.scattering.multipole.moments.size(); ++i)
        multipole_index_.push_back(params_.scattering.multipole.moments[i]);
    qvector_ = q;
    NM = multipole_index_.size();
}

void MPSphereScatterDevice::stage_data() {
    DataStagerByFrame data_stager(sample_, *allcomm_, *partitioncomm_, timer_, be_, ctx_, params_);
    data_stager.stage(SGPU_REPR_SPHERICAL);  // set_representation(SPHERICAL) (:53)
    factors_.assign(NA, 0.0);
}

MPSphereScatterDevice::~MPSphereScatterDevice() {
    if (d_amp_) be_.device_free(d_amp_);
}

// up to `nq` |q| values starting at current_vector_: fills batch_atfinal_/batch_afinal_/batch_a2final_
void MPSphereScatterDevice::compute_batch(size_t nq) {
    timer_.start("sd:c:init");
    std::vector<double> qlens(nq), fb(nq * NA);
    for (size_t i = 0; i < nq; i++) {
        CartesianCoor3D q = vectors_[current_vector_ + i];
        if (i == 0) init_moments(q);
        qlens[i] = q.length();
        sample_.factors(qlens[i], &fb[i * NA]);  // scatterfactors.update(q) per |q|
    }
    ck(be_.set_factors_batch(ctx_, fb.data(), nq, NA), "sgpu_set_factors_batch");
    timer_.stop("sd:c:init");
    const int dsp = dsp_type_code();
    dsp_method_code();
    for (auto &mm : multipole_index_)
        if (labs(mm.second) > mm.first)  // multipole_scatter_device.cpp:459-465
            throw Error("Combination of Major and minor moment not allowed: l=" + std::to_string(mm.first) + ", m" +
                        std::to_string(mm.second));
    std::vector<long> lm(2 * NM);
    for (size_t i = 0; i < NM; i++) {
        lm[2 * i] = multipole_index_[i].first;
        lm[2 * i + 1] = multipole_index_[i].second;
    }
    const size_t amp_len = nq * NM * NF * 2;
    if (amp_len > amp_cap_) {
        if (d_amp_) be_.device_free(d_amp_);
        d_amp_ = nullptr;
        void *pp = nullptr;
        if (be_.device_alloc(&pp, amp_len * sizeof(double))) throw Error("device allocation of the amplitude buffer failed");
        d_amp_ = static_cast<double *>(pp);
        amp_cap_ = amp_len;
    }
    // atom decomposition inside the partition: rank r sums over DivAssignment(NNPP, r, NA) atoms
    DivAssignment mine(partitioncomm_->size(), partitioncomm_->rank(), NA);
    timer_.start("sd:c:block");
    ck(be_.mpsphere_amplitudes(ctx_, qlens.data(), nq, lm.data(), NM, mine.offset(), mine.size(), d_amp_),
       "sgpu_mpsphere_amplitudes");
    timer_.stop("sd:c:block");
    timer_.start("sd:c:wait");
    ck(be_.synchronize(ctx_), "sgpu_synchronize");
    timer_.stop("sd:c:wait");
    timer_.start("sd:c:reduce");
    if (partitioncomm_->size() > 1) partitioncomm_->allreduce_sum(d_amp_, amp_len);
    timer_.stop("sd:c:reduce");
    size_t plen = 0;
    ck(be_.partial_len(ctx_, dsp, &plen), "sgpu_partial_len");
    if (nq * plen > partial_cap_) {
        if (d_partial_) be_.device_free(d_partial_);
        d_partial_ = nullptr;
        void *pp = nullptr;
        if (be_.device_alloc(&pp, nq * plen * sizeof(double))) throw Error("device allocation of the partial buffer failed");
        d_partial_ = static_cast<double *>(pp);
        partial_cap_ = nq * plen;
    }
    timer_.start("sd:c:b:dspstore");
    ck(be_.mpsphere_dsp_partial(ctx_, d_amp_, nq, NM, dsp, d_partial_), "sgpu_mpsphere_dsp_partial");
    batch_atfinal_.assign(nq, std::vector<double>(2 * NF));
    batch_afinal_.assign(nq, 0.0);
    batch_a2final_.assign(nq, 0.0);
    const double factor = 1.0 / (4 * M_PI);  // :395-400
    for (size_t i = 0; i < nq; i++) {
        double af[2], a2f[2];
        ck(be_.finalize(ctx_, d_partial_ + i * plen, dsp, dsp_method_code(), factor, batch_atfinal_[i].data(), af, a2f),
           "sgpu_finalize");
        batch_afinal_[i] = std::complex<double>(af[0], af[1]);
        batch_a2final_[i] = std::complex<double>(a2f[0], a2f[1]);
    }
    timer_.stop("sd:c:b:dspstore");
}

void MPSphereScatterDevice::compute() {
    compute_batch(1);
    atfinal_ = batch_atfinal_[0];
    afinal_ = batch_afinal_[0];
    a2final_ = batch_a2final_[0];
}

void MPSphereScatterDevice::runner() {
    const size_t BATCH = 8;  // |q| values per pass of the batched multipole kernel (multipole_batch_max, multipole.cu; 16 per
                             // pass measured 3 % faster on config 4: not worth a second instantiation)
    while (status() == 0) {
        const size_t nq = std::min(BATCH, vectors_.size() - current_vector_);
        timer_.start("sd:compute");
        compute_batch(nq);
        timer_.stop("sd:compute");
        for (size_t i = 0; i < nq; i++) {
            atfinal_ = batch_atfinal_[i];
            afinal_ = batch_afinal_[i];
            a2final_ = batch_a2final_[i];
            timer_.start("sd:write");
            write();
            timer_.stop("sd:write");
            next();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// MPCylinderScatterDevice (multipole_scatter_device.cpp:504-985)
// ---------------------------------------------------------------------------------------------------------------
MPCylinderScatterDevice::~MPCylinderScatterDevice() {
    if (d_amp_) be_.device_free(d_amp_);
}

void MPCylinderScatterDevice::stage_data() {
    DataStagerByFrame data_stager(sample_, *allcomm_, *partitioncomm_, timer_, be_, ctx_, params_);
    data_stager.stage(SGPU_REPR_CARTESIAN);
    // sample_.coordinate_sets.set_representation(CYLINDRICAL) (:524): CylindricalCoordinateSet around the orientation axis
    const CartesianCoor3D &o = params_.scattering.axis;
    const double axis[3] = {o.x, o.y, o.z};
    ck(be_.frames_to_cylindrical(ctx_, axis), "sgpu_frames_to_cylindrical");
    factors_.assign(NA, 0.0);
}

void MPCylinderScatterDevice::compute() {
    CartesianCoor3D q = vectors_[current_vector_];
    timer_.start("sd:c:init");
    multipole_index_ = params_.scattering.multipole.moments;  // init_moments (:690-700)
    NM = multipole_index_.size();
    sample_.factors(q.length(), factors_.data());  // scatterfactors.update(q)
    ck(be_.set_factors(ctx_, factors_.data(), NA), "sgpu_set_factors");
    std::vector<long> lm(2 * NM);
    for (size_t i = 0; i < NM; i++) {
        lm[2 * i] = multipole_index_[i].first;
        lm[2 * i + 1] = multipole_index_[i].second;
    }
    const int dsp = dsp_type_code();
    dsp_method_code();
    const size_t amp_len = NM * NF * 2;
    if (amp_len > amp_cap_) {
        if (d_amp_) be_.device_free(d_amp_);
        d_amp_ = nullptr;
        void *pp = nullptr;
        if (be_.device_alloc(&pp, amp_len * sizeof(double))) throw Error("device allocation of the amplitude buffer failed");
        d_amp_ = static_cast<double *>(pp);
        amp_cap_ = amp_len;
    }
    timer_.stop("sd:c:init");
    const CartesianCoor3D &o = params_.scattering.axis;
    const double axis[3] = {o.x, o.y, o.z}, qv[3] = {q.x, q.y, q.z};
    // atom decomposition inside the partition: rank r sums over DivAssignment(NNPP, r, NA) atoms (the reference splits
    // the frames, :913-916, and exchanges timelines; amplitudes are sums over atoms, so this needs one all-reduce)
    DivAssignment mine(partitioncomm_->size(), partitioncomm_->rank(), NA);
    timer_.start("sd:c:block");
    ck(be_.mpcylinder_amplitudes(ctx_, qv, axis, lm.data(), NM, mine.offset(), mine.size(), d_amp_), "sgpu_mpcylinder_amplitudes");
    timer_.stop("sd:c:block");
    timer_.start("sd:c:wait");
    ck(be_.synchronize(ctx_), "sgpu_synchronize");
    timer_.stop("sd:c:wait");
    timer_.start("sd:c:reduce");
    if (partitioncomm_->size() > 1) partitioncomm_->allreduce_sum(d_amp_, amp_len);
    timer_.stop("sd:c:reduce");
    double *partial = partial_buffer(dsp);
    timer_.start("sd:c:b:dspstore");
    ck(be_.mpsphere_dsp_partial(ctx_, d_amp_, 1, NM, dsp, partial), "sgpu_mpsphere_dsp_partial");
    double af[2], a2f[2];
    const double factor = 1.0 / (2 * M_PI);  // :866
    ck(be_.finalize(ctx_, partial, dsp, dsp_method_code(), factor, atfinal_.data(), af, a2f), "sgpu_finalize");
    afinal_ = std::complex<double>(af[0], af[1]);
    a2final_ = std::complex<double>(a2f[0], a2f[1]);
    timer_.stop("sd:c:b:dspstore");
}

// ---------------------------------------------------------------------------------------------------------------
// ScatterDeviceFactory (scatter_device_factory.cpp:23-210)
// ---------------------------------------------------------------------------------------------------------------
IScatterDevice *ScatterDeviceFactory::create(std::shared_ptr<ICommunicator> scatter_comm, Sample &sample,
                                             IResultSink *sink, std::vector<CartesianCoor3D> &qvectors,
                                             const Params &params, const SgpuBackend &be, sgpu_ctx *ctx) {
    size_t NN = scatter_comm->size();
    size_t NF = sample.NF;
    size_t NA = sample.NA;
    size_t NQ = qvectors.size();
    if (NF < 1) throw Error("No frames available. Aborting");
    if (NA < 1) throw Error("No atoms available. Aborting");
    if (NQ < 1) throw Error("No qvectors left to compute. Aborting");

    const std::string &stype = params.scattering.type;
    size_t NAF = NA;
    size_t ELBYTESIZE;
    if (stype == "self") {
        NAF = NA;
        ELBYTESIZE = NF * 3 * sizeof(float);
    } else if (stype == "all") {
        NAF = NF;
        ELBYTESIZE = NA * 3 * sizeof(float);
    } else {
        throw Error("Scattering Interference type not understood. Must be 'self' or 'all'.");
    }
    // every rank evaluates the (deterministic) plan; the reference computes it on rank 0 and broadcasts (:95-102)
    // a self run whose atoms stream through the GPU in waves is not constrained by the coordinate budget
    const size_t plan_limit = (stype == "self" && params.limits.stage_stream) ? (size_t)-1 : params.limits.stage_memory_data;
    // multipole devices and the vector-sharded coherent path keep every frame on every rank of a partition: the plan must
    // not count on a larger partition to fit the coordinate budget (it would pass here and fail at stage time)
    const bool replicated = stype == "all" && (params.scattering.orientation_type == "multipole" ||
                                               params.limits.coherent_sharding == "vectors");
    DecompositionPlan dplan(NN, NQ, NAF, ELBYTESIZE, plan_limit, params.limits.decomposition, replicated);
    size_t partitions = dplan.partitions();
    size_t partitionsize = dplan.partitionsize();

    size_t allcommsize = partitions * partitionsize;
    int allcommflag = scatter_comm->rank() < allcommsize ? 1 : 0;
    std::shared_ptr<ICommunicator> all_comm = scatter_comm->split(allcommflag);
    // The reference splits all_comm by partitionID (:117-120).  Here BOTH splits are collectives over scatter_comm, with a
    // colour of their own for the ranks the plan leaves spare: a communicator whose split is job-wide under the hood
    // (torch.distributed new_group, ncclCommSplit) would otherwise wait forever for the spare ranks, which return below.
    // Ranks keep their order, and a partition's ranks are contiguous, so partition_comm is the same group either way.
    size_t partitionID = allcommflag ? (all_comm->rank() * partitions) / allcommsize : partitions;
    std::shared_ptr<ICommunicator> partition_comm = scatter_comm->split((int)partitionID);
    if (allcommflag == 0) return nullptr;

    DivAssignment qindex_assignment(partitions, partitionID, qvectors.size());
    std::vector<CartesianCoor3D> thispartition_QIV;
    for (size_t i = 0; i < qindex_assignment.size(); i++) thispartition_QIV.push_back(qvectors[qindex_assignment[i]]);

    IScatterDevice *p_ScatterDevice = nullptr;
    if (stype == "self") {
        p_ScatterDevice = new SelfVectorsScatterDevice(all_comm, partition_comm, sample, thispartition_QIV, NAF, sink,
                                                       params, be, ctx);
    } else {
        const std::string &otype = params.scattering.orientation_type;
        if (otype == "vectors" || otype == "none") {
            p_ScatterDevice = new AllVectorsScatterDevice(all_comm, partition_comm, sample, thispartition_QIV, NAF, sink,
                                                          params, be, ctx);
        } else if (otype == "multipole") {
            if (params.scattering.multipole.type == "sphere") {
                p_ScatterDevice = new MPSphereScatterDevice(all_comm, partition_comm, sample, thispartition_QIV, NAF,
                                                            sink, params, be, ctx);
            } else if (params.scattering.multipole.type == "cylinder") {
                p_ScatterDevice = new MPCylinderScatterDevice(all_comm, partition_comm, sample, thispartition_QIV, NAF,
                                                              sink, params, be, ctx);
            } else {
                throw Error("scattering.average.orientation.multipole.type not understood: " +
                            params.scattering.multipole.type);
            }
        }
    }
    if (p_ScatterDevice == nullptr) throw Error("Error initializing ScatterDevice");
    return p_ScatterDevice;
}

}  // namespace sassena
