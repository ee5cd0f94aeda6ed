// ref_devices_wrap.cpp — runs the REFERENCE's own scatter devices (src/scatter_devices/{abstract_scatter_device,
// abstract_vectors_scatter_device,all_vectors_scatter_device,self_vectors_scatter_device,multipole_scatter_device}.cpp) with its own stagers
// (src/stager/data_stager.cpp) and DSP (src/math/smath.cpp), compiled where they lie for ONE MPI rank over the shims in
// oracle/shim (communicator, worker threads, Params values, Sample served from arrays, factors and result sink as callbacks,
// FFTW3 API over the oracle's DFT).  Test infrastructure: pins the oracle's compute_all_vectors / compute_self_vectors —
// amplitude loop, alignpad, dsp, store, final scaling, init_subvectors, stager narrowing — against the reference's code.
#include <complex>
#include <cstring>
#include <string>
#include <vector>

#include "control.hpp"
#include "report/timer.hpp"
#include "sample.hpp"
#include "scatter_devices/all_vectors_scatter_device.hpp"
#include "scatter_devices/multipole_scatter_device.hpp"
#include "scatter_devices/self_vectors_scatter_device.hpp"

namespace {
// (the reference's device destructors are protected: devices come from its factory and live until exit)
struct AllDev : AllVectorsScatterDevice {
    using AllVectorsScatterDevice::AllVectorsScatterDevice;
    ~AllDev() {}
};
struct SelfDev : SelfVectorsScatterDevice {
    using SelfVectorsScatterDevice::SelfVectorsScatterDevice;
    ~SelfDev() {}
};
struct MPSphereDev : MPSphereScatterDevice {
    using MPSphereScatterDevice::MPSphereScatterDevice;
    ~MPSphereDev() {}
};
struct MPCylinderDev : MPCylinderScatterDevice {
    using MPCylinderScatterDevice::MPCylinderScatterDevice;
    ~MPCylinderDev() {}
};
struct Out {
    double *fqt, *fq, *fq2, *qout;
    size_t NF, count;
};
void on_write(void *user, const double q[3], const double *fqt, size_t NF, const double fq[2], const double fq2[2]) {
    Out *o = static_cast<Out *>(user);
    std::memcpy(o->fqt + o->count * 2 * NF, fqt, sizeof(double) * 2 * NF);
    std::memcpy(o->fq + 2 * o->count, fq, sizeof(double) * 2);
    std::memcpy(o->fq2 + 2 * o->count, fq2, sizeof(double) * 2);
    std::memcpy(o->qout + 3 * o->count, q, sizeof(double) * 3);
    o->count++;
}
struct Factors {
    const double *b;
};
void on_factors(void *user, double, double *b, size_t n) { std::memcpy(b, static_cast<Factors *>(user)->b, sizeof(double) * n); }
}  // namespace

extern "C" {
// kind: 0 = AllVectorsScatterDevice, 1 = SelfVectorsScatterDevice.  frames: float [NF][NA][3] (the target selection).
// orient: NM unit vectors (vectors_type "sphere" | "file" | "cylinder"; NM = 0: no averaging).  Returns the number of q-vectors
// written; outputs in the order written: fqt [NQ][NF][2], fq [NQ][2], fq2 [NQ][2], qout [NQ][3].
size_t ref_scatter_run(int kind, const float *frames, size_t NF, size_t NA, const double *b, const double *qvectors, size_t NQ,
                       const char *vectors_type, const double *orient, size_t NM, const double axis[3], const char *dsp_type,
                       const char *dsp_method, size_t threads, double *fqt, double *fq, double *fq2, double *qout) {
    Params *p = Params::Inst();
    p->scattering.dsp.type = dsp_type;
    p->scattering.dsp.method = dsp_method;
    p->scattering.average.orientation.vectors.clear();
    p->scattering.average.orientation.vectors.type = vectors_type;
    for (size_t i = 0; i < NM; i++)
        p->scattering.average.orientation.vectors.push_back(CartesianCoor3D(orient[3 * i], orient[3 * i + 1], orient[3 * i + 2]));
    p->scattering.average.orientation.axis = CartesianCoor3D(axis[0], axis[1], axis[2]);
    p->limits.computation.threads = threads;
    p->stager.target = "system";
    p->stager.dump = false;

    Sample sample;
    ShimRangeSelection system(NA);
    sample.atoms.selections["system"] = &system;
    sample.coordinate_sets.shim_set(frames, NF, NA, p->scattering.average.orientation.axis);

    {
        std::lock_guard<std::mutex> l(ShimTimerTable::Inst().m);
        ShimTimerTable::Inst().sum.clear();
    }
    Factors fac = {b};
    ShimFactorSource::Inst().cb = on_factors;
    ShimFactorSource::Inst().user = &fac;
    Out out = {fqt, fq, fq2, qout, NF, 0};
    ShimWriterSink::Inst().cb = on_write;
    ShimWriterSink::Inst().user = &out;

    std::vector<CartesianCoor3D> vectors;
    for (size_t i = 0; i < NQ; i++) vectors.push_back(CartesianCoor3D(qvectors[3 * i], qvectors[3 * i + 1], qvectors[3 * i + 2]));
    boost::mpi::communicator comm;
    boost::asio::ip::tcp::endpoint ep;
    if (kind == 0) {
        AllDev dev(comm, comm, sample, vectors, NF, ep, ep);
        dev.run();
    } else {
        SelfDev dev(comm, comm, sample, vectors, NA, ep, ep);
        dev.run();
    }
    ShimWriterSink::Inst().cb = nullptr;
    ShimFactorSource::Inst().cb = nullptr;
    return out.count;
}

// The same devices on `nranks` ranks of ONE partition, every rank a thread over the shared-memory communicator of
// shim/boost/mpi.hpp: the reference's own multi-rank code runs -- DivAssignment of the frames + all_to_all + alignpad for the
// coherent device (all_vectors_scatter_device.cpp:169-207,291-315), ModAssignment of the atoms + the staged transposition of
// DataStagerByAtom for the self device (data_stager.cpp:249-338), the reductions to partition rank 0 (:335-343).  Partition
// rank 0 writes.  Arguments and outputs as ref_scatter_run; `threads` worker threads per rank.
size_t ref_scatter_run_ranks(int kind, int nranks, const float *frames, size_t NF, size_t NA, const double *b, const double *qvectors,
                             size_t NQ, const char *vectors_type, const double *orient, size_t NM, const double axis[3],
                             const char *dsp_type, const char *dsp_method, size_t threads, double *fqt, double *fq, double *fq2,
                             double *qout) {
    Params *p = Params::Inst();
    p->scattering.dsp.type = dsp_type;
    p->scattering.dsp.method = dsp_method;
    p->scattering.average.orientation.vectors.clear();
    p->scattering.average.orientation.vectors.type = vectors_type;
    for (size_t i = 0; i < NM; i++)
        p->scattering.average.orientation.vectors.push_back(CartesianCoor3D(orient[3 * i], orient[3 * i + 1], orient[3 * i + 2]));
    p->scattering.average.orientation.axis = CartesianCoor3D(axis[0], axis[1], axis[2]);
    p->limits.computation.threads = threads;
    p->stager.target = "system";
    p->stager.dump = false;
    {
        std::lock_guard<std::mutex> l(ShimTimerTable::Inst().m);
        ShimTimerTable::Inst().sum.clear();
    }
    Factors fac = {b};
    ShimFactorSource::Inst().cb = on_factors;
    ShimFactorSource::Inst().user = &fac;
    Out out = {fqt, fq, fq2, qout, NF, 0};
    ShimWriterSink::Inst().cb = on_write;
    ShimWriterSink::Inst().user = &out;
    std::vector<CartesianCoor3D> vectors;
    for (size_t i = 0; i < NQ; i++) vectors.push_back(CartesianCoor3D(qvectors[3 * i], qvectors[3 * i + 1], qvectors[3 * i + 2]));

    std::shared_ptr<boost::mpi::shim_detail::World> world = std::make_shared<boost::mpi::shim_detail::World>(nranks);
    std::vector<std::thread> ranks;
    std::vector<int> failed((size_t)nranks, 0);
    for (int r = 0; r < nranks; r++) {
        ranks.emplace_back([&, r]() {
            try {
                Sample sample;
                ShimRangeSelection system(NA);
                sample.atoms.selections["system"] = &system;
                sample.coordinate_sets.shim_set(frames, NF, NA, Params::Inst()->scattering.average.orientation.axis);
                boost::mpi::communicator allcomm(world, r), partitioncomm(world, r);
                boost::asio::ip::tcp::endpoint ep;
                if (kind == 0) {
                    AllDev dev(allcomm, partitioncomm, sample, vectors, NF, ep, ep);
                    dev.run();
                } else {
                    SelfDev dev(allcomm, partitioncomm, sample, vectors, NA, ep, ep);
                    dev.run();
                }
            } catch (...) {
                failed[(size_t)r] = 1;
            }
        });
    }
    for (auto &t : ranks) t.join();
    ShimWriterSink::Inst().cb = nullptr;
    ShimFactorSource::Inst().cb = nullptr;
    for (int f : failed)
        if (f) return (size_t)-1;
    return out.count;
}

// The reference's multipole devices (src/scatter_devices/multipole_scatter_device.cpp; Boost.Math's three special functions
// served by the oracle's restatements, see shim/boost/math/special_functions.hpp).  kind: 2 = MPSphereScatterDevice,
// 3 = MPCylinderScatterDevice.  frames: float [NF][NA][3] cartesian (the device asks the sample for its own representation),
// moments: NMOM pairs (l, m).  Outputs as ref_scatter_run.
size_t ref_multipole_run(int kind, const float *frames, size_t NF, size_t NA, const double *b, const double *qvectors, size_t NQ,
                         const long *moments, size_t NMOM, const double axis[3], const char *dsp_type, const char *dsp_method,
                         size_t threads, double *fqt, double *fq, double *fq2, double *qout) {
    Params *p = Params::Inst();
    p->scattering.dsp.type = dsp_type;
    p->scattering.dsp.method = dsp_method;
    p->scattering.average.orientation.multipole.moments.clear();
    for (size_t i = 0; i < NMOM; i++)
        p->scattering.average.orientation.multipole.moments.push_back(std::make_pair(moments[2 * i], moments[2 * i + 1]));
    p->scattering.average.orientation.axis = CartesianCoor3D(axis[0], axis[1], axis[2]);
    p->limits.computation.threads = threads;
    p->stager.target = "system";
    p->stager.dump = false;

    Sample sample;
    ShimRangeSelection system(NA);
    sample.atoms.selections["system"] = &system;
    sample.coordinate_sets.shim_set(frames, NF, NA, p->scattering.average.orientation.axis);
    {
        std::lock_guard<std::mutex> l(ShimTimerTable::Inst().m);
        ShimTimerTable::Inst().sum.clear();
    }
    Factors fac = {b};
    ShimFactorSource::Inst().cb = on_factors;
    ShimFactorSource::Inst().user = &fac;
    Out out = {fqt, fq, fq2, qout, NF, 0};
    ShimWriterSink::Inst().cb = on_write;
    ShimWriterSink::Inst().user = &out;

    std::vector<CartesianCoor3D> vectors;
    for (size_t i = 0; i < NQ; i++) vectors.push_back(CartesianCoor3D(qvectors[3 * i], qvectors[3 * i + 1], qvectors[3 * i + 2]));
    boost::mpi::communicator comm;
    boost::asio::ip::tcp::endpoint ep;
    if (kind == 2) {
        MPSphereDev dev(comm, comm, sample, vectors, NF, ep, ep);
        dev.run();
    } else {
        MPCylinderDev dev(comm, comm, sample, vectors, NF, ep, ep);
        dev.run();
    }
    ShimWriterSink::Inst().cb = nullptr;
    ShimFactorSource::Inst().cb = nullptr;
    return out.count;
}

// wall-clock seconds the last ref_scatter_run spent under a timer key of the reference ("sd:stage", "sd:runner", "sd:compute", ...)
double ref_timer_seconds(const char *key) {
    std::lock_guard<std::mutex> l(ShimTimerTable::Inst().m);
    std::map<std::string, double>::iterator it = ShimTimerTable::Inst().sum.find(key);
    return it == ShimTimerTable::Inst().sum.end() ? 0.0 : it->second;
}
}
