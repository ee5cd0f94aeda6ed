/* boost/mpi.hpp — SHIM: a one-rank communicator.  The reference's scatter devices are built here for a single MPI rank
 * (partition size 1), so collectives are copies. */
#ifndef ORACLE_SHIM_BOOST_MPI_HPP
#define ORACLE_SHIM_BOOST_MPI_HPP
#include <cstddef>
#include <cstring>
#include <functional>
#include <vector>
#include <boost/lexical_cast.hpp>  /* (the real header pulls it in; data_stager.cpp relies on that) */
namespace boost { namespace mpi {
template <class T> struct maximum { const T &operator()(const T &a, const T &b) const { return a < b ? b : a; } };
template <class T> struct minimum { const T &operator()(const T &a, const T &b) const { return a < b ? a : b; } };
class communicator {
   public:
    int rank() const { return 0; }
    int size() const { return 1; }
    void barrier() const {}
    communicator split(int) const { return *this; }
};
class environment {
   public:
    environment() {}
    template <class A, class B> environment(A &, B &) {}
};
template <class T> void all_to_all(const communicator &, const T *in, int n, T *out) { std::memcpy(out, in, sizeof(T) * n); }
template <class T, class Op> void reduce(const communicator &, const T *in, int n, T *out, Op, int) { if (out) std::memcpy(out, in, sizeof(T) * n); }
template <class T, class Op> void reduce(const communicator &, const T &in, T &out, Op, int) { out = in; }
template <class T, class Op> void all_reduce(const communicator &, const T *in, int n, T *out, Op) { std::memcpy(out, in, sizeof(T) * n); }
template <class T, class Op> void all_reduce(const communicator &, const T &in, T &out, Op) { out = in; }
template <class T> void broadcast(const communicator &, T *, int, int) {}
template <class T> void broadcast(const communicator &, T &, int) {}
template <class T> void gather(const communicator &, const T &in, std::vector<T> &out, int) { out.assign(1, in); }
template <class T> void all_gather(const communicator &, const T &in, std::vector<T> &out) { out.assign(1, in); }
}}  // namespace boost::mpi
#endif
