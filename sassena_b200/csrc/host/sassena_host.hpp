// sassena_host.hpp — host-side mirror of the reference's operator interface for the scattering hot path.
//
// Same names, argument meaning and control flow as the reference (benlabs/sassena v1.4.2):
//   DivAssignment / ModAssignment      src/decomposition/assignment.cpp:27-132
//   DecompositionPlan                  src/decomposition/decomposition_plan.cpp:29-156
//   orientation / scan / moment generators   src/control/parameters.cpp:930-1189
//   CartesianVectorBase                src/math/coor3d.cpp:278-304
//   DataStagerByFrame / ByAtom         src/stager/data_stager.cpp:39-349
//   IScatterDevice, AbstractScatterDevice, AbstractVectorsScatterDevice, AllVectorsScatterDevice,
//   SelfVectorsScatterDevice, MPSphereScatterDevice, ScatterDeviceFactory    src/scatter_devices/*
// What changed is what sits underneath: compute() drives the CUDA kernels through the C-ABI
// (include/sassena_b200.h), a "rank" is one GPU, MPI collectives become one NCCL all-reduce of the packed
// partial per |q|, and the boost::asio result service becomes an IResultSink callback.
// Errors: the reference does Err::write(msg) + throw; here every such site throws sassena::Error(msg),
// which the C wrapper (sass_capi.cpp) turns into a return code.
#pragma once

#include <complex>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <cmath>
#include <map>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../../include/sassena_host.h"

namespace sassena {

struct Error : std::runtime_error {
    explicit Error(const std::string &m) : std::runtime_error(m) {}
};
// reference: sassena::terminate_request (include/exceptions/exceptions.hpp:26-29), thrown when ram_check fails
struct terminate_request : Error {
    terminate_request() : Error("terminate_request: memory limits exceeded") {}
};

// ---------------------------------------------------------------------------------------------------------------
// Boost.Random streams the reference draws from (parameters.cpp:948-959,1002-1013; motion_walker.cpp)
// ---------------------------------------------------------------------------------------------------------------
// boost::normal_distribution<double> over boost::mt19937 as Boost 1.4x implements it: Box-Muller with one cached value,
// uniform_01 = x / 2^32.  std::mt19937 is the same engine and seeding as boost::mt19937.
// BEST EFFORT: the stream of later Boost versions differs; parity runs pass explicit vectors.
struct BoostNormal {
    std::mt19937 rng;
    bool valid = false;
    double r1 = 0, cached_rho = 0;
    explicit BoostNormal(uint32_t seed) : rng(seed) {}
    double operator()() {
        if (!valid) {
            r1 = rng() / 4294967296.0;
            double r2 = rng() / 4294967296.0;
            cached_rho = std::sqrt(-2.0 * std::log(1.0 - r2));
            valid = true;
            return cached_rho * std::cos(2 * M_PI * r1);
        }
        valid = false;
        return cached_rho * std::sin(2 * M_PI * r1);
    }
};
// boost::uniform_on_sphere<double>(dim): dim normal variates, normalised
struct UniformOnSphere {
    BoostNormal normal;
    int dim;
    UniformOnSphere(uint32_t seed, int d) : normal(seed), dim(d) {}
    std::vector<double> operator()() {
        std::vector<double> v(dim);
        double sq = 0;
        for (int d = 0; d < dim; d++) {
            v[d] = normal();
            sq += v[d] * v[d];
        }
        double inv = 1.0 / std::sqrt(sq);
        for (auto &c : v) c *= inv;
        return v;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// math/coor3d
// ---------------------------------------------------------------------------------------------------------------
struct CartesianCoor3D {
    double x = 0, y = 0, z = 0;
    CartesianCoor3D() = default;
    CartesianCoor3D(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
    double length() const;  // coor3d.cpp:58-60
    CartesianCoor3D operator+(const CartesianCoor3D &o) const { return {x + o.x, y + o.y, z + o.z}; }
    CartesianCoor3D operator-(const CartesianCoor3D &o) const { return {x - o.x, y - o.y, z - o.z}; }
    double operator*(const CartesianCoor3D &o) const { return x * o.x + y * o.y + z * o.z; }
    CartesianCoor3D cross_product(const CartesianCoor3D &o) const {
        return {y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x};
    }
};
inline CartesianCoor3D operator*(double l, const CartesianCoor3D &c) { return {l * c.x, l * c.y, l * c.z}; }
inline CartesianCoor3D operator/(const CartesianCoor3D &c, double l) { return {c.x / l, c.y / l, c.z / l}; }

// orthonormal base (e_r, e_phi, e_z) built "out of thin air" from an axis, coor3d.cpp:278-304
class CartesianVectorBase {
    std::vector<CartesianCoor3D> base_;

   public:
    explicit CartesianVectorBase(CartesianCoor3D axis);
    CartesianCoor3D &operator[](size_t i) { return base_.at(i); }
    CartesianCoor3D project(CartesianCoor3D vec);
};

// ---------------------------------------------------------------------------------------------------------------
// decomposition
// ---------------------------------------------------------------------------------------------------------------
class DivAssignment {
    size_t NN_, rank_, NAF_, offset_, size_;

   public:
    DivAssignment(size_t NN, size_t rank, size_t NAF);
    size_t operator[](size_t index) const;
    size_t size() const { return size_; }
    size_t offset() const { return offset_; }
    size_t max() const;
    bool contains(size_t i) const;
    size_t index(size_t i) const;
};

class ModAssignment {
    size_t NN_, rank_, NAF_, offset_, size_;

   public:
    ModAssignment(size_t NN, size_t rank, size_t NAF);
    size_t operator[](size_t index) const;
    size_t size() const { return size_; }
    size_t offset() const { return offset_; }
    size_t max() const;
    bool contains(size_t i) const;
    size_t index(size_t i) const;
};

class DecompositionParameters {
    size_t m_penalty, m_NAFcycles, m_NQcycles, m_NNpP, m_NN, m_NQ, m_NAF, m_NP, m_elbytesize, m_nbytesize;

   public:
    DecompositionParameters(size_t NN, size_t NQ, size_t NAF, size_t NNpP, size_t elbytesize, bool replicated = false);
    size_t penalty() const { return m_penalty; }
    size_t nbytesize() const { return m_nbytesize; }
    size_t get_NN() const { return m_NN; }
    size_t get_NQ() const { return m_NQ; }
    size_t get_NAF() const { return m_NAF; }
    size_t get_NP() const { return m_NP; }
    size_t get_NNpP() const { return m_NNpP; }
    size_t get_NAFcycles() const { return m_NAFcycles; }
    size_t get_NQcycles() const { return m_NQcycles; }
};

struct DecompositionLimits {  // limits.decomposition.* (parameters.cpp:634-636)
    double utilization = 0.95;
    bool partitions_automatic = true;
    size_t partitions_size = 1;
};

class DecompositionPlan {
    std::unique_ptr<DecompositionParameters> p_dp_best;

   public:
    // replicated: every rank of a partition stages all frames (multipole devices, vector-sharded coherent path)
    DecompositionPlan(size_t nn, size_t nq, size_t naf, size_t elbytesize, size_t nmaxbytesize,
                      const DecompositionLimits &lim = DecompositionLimits(), bool replicated = false);
    size_t partitions() const { return p_dp_best->get_NP(); }
    size_t partitionsize() const { return p_dp_best->get_NNpP(); }
    size_t penalty() const { return p_dp_best->penalty(); }
    double utilization() const;
};

// ---------------------------------------------------------------------------------------------------------------
// control: the part of Params the hot path reads (src/control/parameters.cpp:372-397, 612-636)
// ---------------------------------------------------------------------------------------------------------------
struct ScatteringVectorsScan {  // scattering.vectors.scans.scan (parameters.cpp:453-484)
    CartesianCoor3D basevector{1, 0, 0};
    double from = 0, to = 1, exponent = 1.0;
    size_t points = 100;
};
// ScatteringVectorsParameters::create_from_scans, parameters.cpp:1125-1189 (powf-rounded fractions)
std::vector<CartesianCoor3D> create_from_scans(const std::vector<ScatteringVectorsScan> &scans);

struct OrientationVectorsParameters {  // scattering.average.orientation.vectors.*
    std::string type = "sphere";       // sphere | cylinder | file
    std::string algorithm = "boost_uniform_on_sphere";
    size_t resolution = 100;
    uint32_t seed = 0;
    std::vector<CartesianCoor3D> vectors;  // filled by create() or given explicitly (type=file rows)
    // ScatteringAverageOrientationVectorsParameters::create, parameters.cpp:930-1034.
    // For type=="file" the rows already in `vectors` are normalised (:932-943).
    void create();
};

struct OrientationMultipoleParameters {  // scattering.average.orientation.multipole.*
    std::string type = "sphere";         // sphere | cylinder
    std::string moments_type;            // resolution | file  (no default in the reference, :1037-1079)
    long resolution = 20;
    std::vector<std::pair<long, long>> moments;
    void create();  // parameters.cpp:1037-1122
};

struct ScatteringParameters {
    std::string type = "all";  // all | self
    std::string dsp_type = "autocorrelate";
    std::string dsp_method = "fftw";
    std::string orientation_type = "none";  // none | vectors | multipole
    CartesianCoor3D axis{0, 0, 1};
    OrientationVectorsParameters vectors;
    OrientationMultipoleParameters multipole;
};

struct LimitsParameters {
    // limits.stage.memory.data re-targeted: bytes of coordinates one GPU may hold (default 150 GB of the 180 GB HBM3e)
    size_t stage_memory_data = (size_t)150 << 30;
    // limits.stage.stream (self scattering): when a rank's share of the trajectory exceeds stage_memory_data the stager
    // streams its atoms through the GPU in waves (wave-outer, |q|-inner; partials are additive over atoms) instead of
    // failing like the reference (data_stager.cpp:194-204).  false restores the reference's error.
    bool stage_stream = true;
    DecompositionLimits decomposition;
    // how the ranks of one partition share a coherent |q|: "frames" = each rank holds a block of frames (the
    // reference's decomposition, all_vectors_scatter_device.cpp:61,248,408), "vectors" = every rank holds all frames
    // and takes a block of the orientation vectors, "auto" = frames unless the sample is too small to amortise the
    // amplitude exchange
    std::string coherent_sharding = "auto";
    // |q|-scan batching of the coherent path: largest number of consecutive q-vectors evaluated in one pass
    // (0 or 1 disables; limits.computation.scan)
    size_t coherent_scan = 56;
    // limits.computation.scan_snap: move the |q| of a batch that is equally spaced to within 1e-6 onto the exact
    // progression (the reference computes scan fractions in float, parameters.cpp:1151, which perturbs them by ~3e-8).
    // Off by default: it changes the q-vectors (and therefore the results, by ~1e-5 relative) with respect to the
    // reference; on, such scans take the plain scan kernel instead of the corrected one and the snapped q-vectors are
    // what is written.
    bool coherent_scan_snap = false;
};

// stager section (parameters.cpp:345-369): dump = write the staged coordinates to `filepath` (data_stager.cpp:91-94,233-238)
struct StagerParameters {
    bool dump = false;
    std::string file = "dump.dcd", filepath = "dump.dcd", format = "dcd", mode = "frames";
};

struct Params {
    ScatteringParameters scattering;
    LimitsParameters limits;
    StagerParameters stager;
};

// ---------------------------------------------------------------------------------------------------------------
// plumbing that replaces boost::mpi::communicator, HDF5WriterClient and the sgpu library binding
// ---------------------------------------------------------------------------------------------------------------
class ICommunicator {
   public:
    virtual ~ICommunicator() {}
    virtual size_t rank() const = 0;
    virtual size_t size() const = 0;
    // sum `n` doubles in device memory over the ranks, result on every rank (NCCL all-reduce, f64)
    virtual void allreduce_sum(double *d_buf, size_t n) = 0;
    virtual void barrier() = 0;
    // boost::mpi::communicator::split(color): ranks with equal color form a new communicator, ordered by old rank
    virtual std::shared_ptr<ICommunicator> split(int color) = 0;
    // the ncclComm_t behind the communicator, if any (lets the library exchange amplitudes on the device streams)
    virtual void *nccl_comm() { return nullptr; }
};

class SingleCommunicator : public ICommunicator {
   public:
    size_t rank() const override { return 0; }
    size_t size() const override { return 1; }
    void allreduce_sum(double *, size_t) override {}
    void barrier() override {}
    std::shared_ptr<ICommunicator> split(int) override { return std::make_shared<SingleCommunicator>(); }
};

// C callbacks (torch.distributed / NCCL live on the caller's side): sass_comm_vtbl, include/sassena_host.h
class CallbackCommunicator : public ICommunicator {
    sass_comm_vtbl v_;
    bool owned_;

   public:
    CallbackCommunicator(const sass_comm_vtbl &v, bool owned) : v_(v), owned_(owned) {}
    ~CallbackCommunicator() override;
    size_t rank() const override { return v_.rank(v_.user); }
    size_t size() const override { return v_.size(v_.user); }
    void allreduce_sum(double *d, size_t n) override;
    void barrier() override;
    std::shared_ptr<ICommunicator> split(int color) override;
    void *nccl_comm() override { return v_.nccl_comm ? v_.nccl_comm(v_.user) : nullptr; }
};

// the C-ABI as a table (sass_backend_vtbl, include/sassena_host.h), so that the host logic can be exercised
// without a GPU in tests (tests/ bind it to the CPU oracle)
using SgpuBackend = sass_backend_vtbl;
const SgpuBackend &default_backend();  // the real library (sgpu_* of this .so)

// Sample: what the hot path needs from sample_ (atoms in stager.target, frames, scattering factors)
struct Sample {
    size_t NA = 0, NF = 0;
    const float *frames = nullptr;  // host [NF][NA][3], cartesian (CoordinateSets::load order, data_stager.cpp:102-118)
    // ScatterFactors::update(q) + get_all() (scatter_factors.cpp:56-78,100): fill b[NA] for |q|
    std::function<void(double ql, double *b)> factors;
};

// HDF5WriterClient::write(qvector, data, NF, data2, data3) stand-in (file_writer_service.cpp:504-531)
class IResultSink {
   public:
    virtual ~IResultSink() {}
    virtual void write(CartesianCoor3D qvector, const double *fqt, size_t NF, std::complex<double> fq,
                       std::complex<double> fq2) = 0;
};

// named wall-clock timers with the reference's keys (sd:stage, sd:compute, sd:c:init, ... report/timer.cpp:35-59)
class Timer {
    std::map<std::string, std::pair<double, size_t>> acc_;  // sum seconds, count
    std::map<std::string, double> start_;

   public:
    void start(const std::string &key);
    void stop(const std::string &key);
    double sum(const std::string &key) const;
    size_t count(const std::string &key) const;
    std::vector<std::string> keys() const;
};

// ---------------------------------------------------------------------------------------------------------------
// stager
// ---------------------------------------------------------------------------------------------------------------
class DataStagerByFrame {  // data_stager.cpp:39-129
    Sample &m_sample;
    ICommunicator &allcomm_, &partitioncomm_;
    Timer &timer_;
    const SgpuBackend &be_;
    sgpu_ctx *ctx_;
    const Params &params_;

   public:
    DataStagerByFrame(Sample &sample, ICommunicator &allcomm, ICommunicator &partitioncomm, Timer &timer,
                      const SgpuBackend &be, sgpu_ctx *ctx, const Params &params);
    void stage(int repr);  // coordinates end up resident on this rank's GPU, frame-major
    // frame decomposition inside the partition (data_stager.cpp:57-118): this rank stages DivAssignment(NNPP, rank, NF)
    void stage_block();
    // stager.dump (data_stager.cpp:131-165): the first partition writes the frames as a DCD file, rank r its DivAssignment block
    void write(const std::string &filename, const std::string &format);
};

class DataStagerByAtom {  // data_stager.cpp:176-349
    Sample &m_sample;
    ICommunicator &allcomm_, &partitioncomm_;
    Timer &timer_;
    const SgpuBackend &be_;
    sgpu_ctx *ctx_;
    const Params &params_;

   public:
    DataStagerByAtom(Sample &sample, ICommunicator &allcomm, ICommunicator &partitioncomm, Timer &timer,
                     const SgpuBackend &be, sgpu_ctx *ctx, const Params &params);
    void stage();  // ModAssignment(partition size, partition rank, NA) atoms, atom-major on the GPU
    // stager.dump (data_stager.cpp:352-391): a DCD file of NA "frames" with NF "atoms" each (the atom-major staging layout),
    // rank r of the first partition writes the timelines of its ModAssignment atoms
    void write(const std::string &filename, const std::string &format);
};

// ---------------------------------------------------------------------------------------------------------------
// scatter devices
// ---------------------------------------------------------------------------------------------------------------
class IScatterDevice {
   protected:
    virtual void runner() = 0;
    virtual size_t status() = 0;
    virtual double progress() = 0;

   public:
    virtual ~IScatterDevice() {}
    virtual Timer &getTimer() = 0;
    virtual void run() = 0;
};

class AbstractScatterDevice : public IScatterDevice {
   protected:
    std::shared_ptr<ICommunicator> allcomm_, partitioncomm_;
    Sample &sample_;
    std::vector<CartesianCoor3D> vectors_;
    size_t current_vector_ = 0;
    IResultSink *p_hdf5writer_;
    const Params &params_;
    const SgpuBackend &be_;
    sgpu_ctx *ctx_ = nullptr;
    bool own_ctx_ = false;

    size_t NN, NF, NA;
    std::vector<double> atfinal_;  // [NF][2]
    std::complex<double> afinal_, a2final_;
    std::vector<double> factors_;  // ScatterFactors::get_all() of this rank's staged atoms
    double *d_partial_ = nullptr;
    size_t partial_cap_ = 0;
    Timer timer_;

    virtual void stage_data() = 0;
    virtual void compute() = 0;
    void next();
    void write();
    void runner() override;
    virtual void print_pre_stage_info() {}
    virtual void print_post_stage_info() {}
    virtual void print_pre_runner_info() {}
    virtual void print_post_runner_info() {}
    virtual bool ram_check();
    size_t status() override;
    double progress() override;

    // helpers shared by the concrete devices
    void ck(int rc, const char *what);
    int dsp_type_code() const;
    int dsp_method_code() const;
    double *partial_buffer(int dsp_type);
    // d_partial: the packed partial to reduce (default: the device's own partial buffer)
    void reduce_and_finalize(int dsp_type, double scale, double *d_partial = nullptr);

   public:
    AbstractScatterDevice(std::shared_ptr<ICommunicator> allcomm, std::shared_ptr<ICommunicator> partitioncomm,
                          Sample &sample, std::vector<CartesianCoor3D> vectors, size_t NAF, IResultSink *sink,
                          const Params &params, const SgpuBackend &be, sgpu_ctx *ctx);
    ~AbstractScatterDevice() override;
    Timer &getTimer() override { return timer_; }
    void run() override;
};

class AbstractVectorsScatterDevice : public AbstractScatterDevice {
   protected:
    std::vector<CartesianCoor3D> subvector_index_;
    size_t NM = 0;
    size_t current_subvector_ = 0;
    double progress() override;
    void init_subvectors(CartesianCoor3D &q);  // abstract_vectors_scatter_device.cpp:112-175

   public:
    using AbstractScatterDevice::AbstractScatterDevice;
    const std::vector<CartesianCoor3D> &subvectors() const { return subvector_index_; }
};

class AllVectorsScatterDevice : public AbstractVectorsScatterDevice {
   protected:
    bool frame_sharded_ = false;
    bool nccl_sharded_ = false;  // the partition communicator is NCCL: exchange and reduction run inside the library
    double *d_amp_ = nullptr;  // complex A[NM][NF] of the frame-sharded path (device)
    size_t amp_cap_ = 0;
    void stage_data() override;
    void compute() override;
    void compute_frame_sharded();
    // |q|-scan batching of the runner loop (abstract_scatter_device.cpp:162-173): consecutive q-vectors whose
    // subvectors are s_n v_m for common directions v_m go to the backend in one call, which evaluates (nearly) equally
    // spaced |q| with two sincos + a recurrence per (atom, direction) instead of one sincos per |q|
    // (include/sassena_b200.h "|q|-scan coherent path")
    void runner() override;
    size_t scan_length(size_t first, std::vector<double> &v, std::vector<double> &s);
    void compute_scan(size_t nq, const std::vector<double> &v, const std::vector<double> &s);
    std::vector<std::vector<double>> batch_atfinal_;
    std::vector<std::complex<double>> batch_afinal_, batch_a2final_;
    size_t scans_ = 0;  // number of batched passes taken (diagnostics / tests)

   public:
    using AbstractVectorsScatterDevice::AbstractVectorsScatterDevice;
    ~AllVectorsScatterDevice() override;
    bool frame_sharded() const { return frame_sharded_; }
    size_t scan_batches() const { return scans_; }
};

class SelfVectorsScatterDevice : public AbstractVectorsScatterDevice {
   protected:
    ModAssignment assignment_;
    // streamed mode (BASELINE config 5): this rank's atoms pass through the GPU in waves of wave_atoms_
    bool streamed_ = false;
    size_t wave_atoms_ = 0, waves_ = 0;
    double *d_acc_ = nullptr;  // [vectors_.size()][partial_len] partials summed over the waves
    float *h_atoms_ = nullptr;  // pinned, [assignment_.size()][NF][3]: this rank's atoms, atom-major (streamed mode)
    void gather_atoms_to_host();
    void stage_data() override;
    void compute() override;
    void runner() override;
    // factors of staged atoms [first, first+count) of assignment_, subvectors of vectors_[current_vector_]: one |q| into d_out
    void compute_partial(size_t first, size_t count, int dsp, double *d_out);

   public:
    ~SelfVectorsScatterDevice() override;
    bool streamed() const { return streamed_; }
    size_t waves() const { return waves_; }
    SelfVectorsScatterDevice(std::shared_ptr<ICommunicator> allcomm, std::shared_ptr<ICommunicator> partitioncomm,
                             Sample &sample, std::vector<CartesianCoor3D> vectors, size_t NAF, IResultSink *sink,
                             const Params &params, const SgpuBackend &be, sgpu_ctx *ctx);
};

class MPSphereScatterDevice : public AbstractScatterDevice {
   protected:
    std::vector<std::pair<long, long>> multipole_index_;
    CartesianCoor3D qvector_;
    size_t NM = 0;
    double *d_amp_ = nullptr;
    size_t amp_cap_ = 0;
    void init_moments(CartesianCoor3D &q);  // multipole_scatter_device.cpp:155-165
    void stage_data() override;
    void compute() override;
    // The runner loop (abstract_scatter_device.cpp:162-173) processes |q| values in batches: Y_lm does not depend on
    // |q|, so one pass over the atoms serves up to 8 |q|; atoms are sharded over the partition's GPUs and the
    // amplitudes A_lm[f] are all-reduced BEFORE the DSP (they are sums over atoms).
    void runner() override;
    void compute_batch(size_t nq);
    std::vector<std::vector<double>> batch_atfinal_;
    std::vector<std::complex<double>> batch_afinal_, batch_a2final_;

   public:
    using AbstractScatterDevice::AbstractScatterDevice;
    ~MPSphereScatterDevice() override;
};

// MPCylinderScatterDevice (multipole_scatter_device.cpp:504-985): frames staged in cylindrical coordinates around
// scattering.average.orientation.axis, one pass over the atoms per |q| serves all moments (orders of J_n); atoms are
// sharded over the partition's GPUs and the amplitudes all-reduced before the DSP, result scaled by 1/(2 pi).
class MPCylinderScatterDevice : public AbstractScatterDevice {
   protected:
    std::vector<std::pair<long, long>> multipole_index_;
    size_t NM = 0;
    double *d_amp_ = nullptr;
    size_t amp_cap_ = 0;
    void stage_data() override;
    void compute() override;

   public:
    using AbstractScatterDevice::AbstractScatterDevice;
    ~MPCylinderScatterDevice() override;
};

class ScatterDeviceFactory {  // scatter_device_factory.cpp:23-210
   public:
    // returns nullptr on spare ranks (scatter_device_factory.cpp:116-120)
    static IScatterDevice *create(std::shared_ptr<ICommunicator> scatter_comm, Sample &sample, IResultSink *sink,
                                  std::vector<CartesianCoor3D> &qvectors, const Params &params,
                                  const SgpuBackend &be = default_backend(), sgpu_ctx *ctx = nullptr);
};

}  // namespace sassena
