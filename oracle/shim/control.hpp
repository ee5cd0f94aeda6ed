/* control.hpp — SHIM.  The reference's Params / Database singletons parse scatter.xml and db.xml (Boost + libxml2).  Here Params
 * is a plain struct holding the values the code built in oracle/_ref reads; the test harness fills them in. */
#ifndef ORACLE_SHIM_CONTROL_HPP
#define ORACLE_SHIM_CONTROL_HPP
#include <cstddef>
#include <string>
#include <utility>
#include <vector>
#include "math/coor3d.hpp"
#include "exceptions/exceptions.hpp"
struct ShimPartitions { bool automatic = true; size_t size = 1; };
struct ShimDecomposition { ShimPartitions partitions; double utilization = 0.95; };
struct ShimCompMemory { size_t scale = 1, result_buffer = (size_t)1 << 40, signal_buffer = (size_t)1 << 40, exchange_buffer = (size_t)1 << 40, alignpad_buffer = (size_t)1 << 40; };
struct ShimComputation { size_t threads = 1; ShimCompMemory memory; };
struct ShimStageMemory { size_t data = (size_t)1 << 40, buffer = (size_t)1 << 30; };
struct ShimStage { ShimStageMemory memory; };
struct ShimMonitor { size_t sampling = 1; };
struct ShimServices { ShimMonitor monitor; };
struct ShimLimits { ShimDecomposition decomposition; ShimComputation computation; ShimStage stage; ShimServices services; };
struct ShimVectors : public std::vector<CartesianCoor3D> { std::string type = "file"; };
struct ShimMoments : public std::vector<std::pair<long, long> > {};
struct ShimMultipole { ShimMoments moments; };
struct ShimOrientation { std::string type = "vectors"; ShimVectors vectors; ShimMultipole multipole; CartesianCoor3D axis = CartesianCoor3D(0, 0, 1); };
struct ShimAverage { ShimOrientation orientation; };
struct ShimDsp { std::string type = "autocorrelate", method = "fftw"; };
struct ShimScattering { ShimDsp dsp; ShimAverage average; };
struct ShimStager { bool dump = false; std::string filepath = "dump.dcd", format = "dcd", target = "system"; };
struct ShimDebugPrint { bool orientations = false; };
struct ShimDebug { ShimDebugPrint print; };
struct ShimDatabase { std::string type = "file", filepath = "db.xml", format = "xml"; };
class Params {
   public:
    ShimLimits limits;
    ShimScattering scattering;
    ShimStager stager;
    ShimDebug debug;
    ShimDatabase database;
    static Params *Inst() {
        static Params p;
        return &p;
    }
};
#endif
