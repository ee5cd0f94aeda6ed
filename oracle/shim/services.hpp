/* services.hpp — SHIM: the result writer and the progress monitor of the reference are TCP clients of services on rank 0
 * (HDF5 + Boost.Asio).  Here write() hands the result to a callback the test harness installs. */
#ifndef ORACLE_SHIM_SERVICES_HPP
#define ORACLE_SHIM_SERVICES_HPP
#include <complex>
#include <cstddef>
#include <vector>
#include <boost/asio.hpp>
#include <fftw3.h>
#include "math/coor3d.hpp"
typedef void (*shim_write_cb)(void *user, const double q[3], const double *fqt, size_t NF, const double fq[2], const double fq2[2]);
struct ShimWriterSink {
    shim_write_cb cb = nullptr;
    void *user = nullptr;
    static ShimWriterSink &Inst() {
        static ShimWriterSink s;
        return s;
    }
};
class HDF5WriterClient {
   public:
    explicit HDF5WriterClient(boost::asio::ip::tcp::endpoint) {}
    void write(CartesianCoor3D qvector, const fftw_complex *data, size_t NF, const std::complex<double> data2, const std::complex<double> data3) {
        const double q[3] = {qvector.x, qvector.y, qvector.z};
        const double a[2] = {data2.real(), data2.imag()}, b[2] = {data3.real(), data3.imag()};
        if (ShimWriterSink::Inst().cb) ShimWriterSink::Inst().cb(ShimWriterSink::Inst().user, q, &data[0][0], NF, a, b);
    }
    void write(CartesianCoor3D qvector, const std::vector<std::complex<double> > &data, const std::complex<double> data2, const std::complex<double> data3) {
        write(qvector, reinterpret_cast<const fftw_complex *>(data.data()), data.size(), data2, data3);
    }
    void flush() {}
};
class MonitorClient {
   public:
    explicit MonitorClient(boost::asio::ip::tcp::endpoint) {}
    void reset_server() {}
    void set_samplingfactor(size_t) {}
    void set_samplingfactor_server(size_t) {}
    void update(size_t, double) {}
};
#endif
