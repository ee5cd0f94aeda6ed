/* boost/mpi.hpp — SHIM: Boost.MPI's communicator and the collectives the reference's scatter devices and stagers call,
 * for ranks that are THREADS of one process.  A default-constructed communicator is a world of one rank (collectives are
 * copies), which is what ref_scatter_run uses; ref_scatter_run_ranks starts one thread per rank over a shared World, so that
 * the reference's multi-rank branches -- the all_to_all + alignpad of the frame decomposition
 * (all_vectors_scatter_device.cpp:169-207), the staged transposition of DataStagerByAtom (data_stager.cpp:249-338), the
 * reductions to partition rank 0 -- run here too.  Every collective publishes the caller's buffer, meets the others at a
 * barrier, copies / reduces in rank order (deterministic) and meets them again before the buffers may be reused. */
#ifndef ORACLE_SHIM_BOOST_MPI_HPP
#define ORACLE_SHIM_BOOST_MPI_HPP
#include <condition_variable>
#include <cstddef>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <vector>
#include <boost/lexical_cast.hpp>  /* (the real header pulls it in; data_stager.cpp relies on that) */
namespace boost { namespace mpi {
template <class T> struct maximum { const T &operator()(const T &a, const T &b) const { return a < b ? b : a; } };
template <class T> struct minimum { const T &operator()(const T &a, const T &b) const { return a < b ? a : b; } };

namespace shim_detail {
struct World {
    int size;
    std::mutex m;
    std::condition_variable cv;
    int waiting = 0;
    unsigned long generation = 0;
    std::vector<const void *> slot;  // one published pointer per rank
    std::vector<long> ivalue;        // one published integer per rank (split colours)
    std::map<int, std::shared_ptr<World> > children;  // worlds under construction by split(), by colour
    explicit World(int n) : size(n), slot(n, nullptr), ivalue(n, 0) {}
    void barrier() {
        if (size == 1) return;
        std::unique_lock<std::mutex> l(m);
        const unsigned long gen = generation;
        if (++waiting == size) {
            waiting = 0;
            generation++;
            cv.notify_all();
        } else {
            cv.wait(l, [&] { return generation != gen; });
        }
    }
};
}  // namespace shim_detail

class communicator {
    std::shared_ptr<shim_detail::World> w_;
    int rank_;

   public:
    communicator() : w_(std::make_shared<shim_detail::World>(1)), rank_(0) {}
    communicator(std::shared_ptr<shim_detail::World> w, int rank) : w_(w), rank_(rank) {}
    int rank() const { return rank_; }
    int size() const { return w_->size; }
    void barrier() const { w_->barrier(); }
    shim_detail::World &world() const { return *w_; }
    // ranks of equal colour form a new communicator, ordered by their rank here
    communicator split(int color) const {
        if (w_->size == 1) return *this;
        w_->ivalue[rank_] = color;
        w_->barrier();
        int newrank = 0, newsize = 0, first = -1;
        for (int r = 0; r < w_->size; r++)
            if (w_->ivalue[r] == color) {
                if (first < 0) first = r;
                if (r < rank_) newrank++;
                newsize++;
            }
        if (first == rank_) {
            std::lock_guard<std::mutex> l(w_->m);
            w_->children[color] = std::make_shared<shim_detail::World>(newsize);
        }
        w_->barrier();
        std::shared_ptr<shim_detail::World> child;
        {
            std::lock_guard<std::mutex> l(w_->m);
            child = w_->children[color];
        }
        w_->barrier();
        if (first == rank_) {
            std::lock_guard<std::mutex> l(w_->m);
            w_->children.erase(color);
        }
        w_->barrier();
        return communicator(child, newrank);
    }
};
class environment {
   public:
    environment() {}
    template <class A, class B> environment(A &, B &) {}
};

// out[s*n .. s*n+n) = rank s's in[rank*n .. rank*n+n)
template <class T> void all_to_all(const communicator &c, const T *in, int n, T *out) {
    if (c.size() == 1) { std::memcpy(out, in, sizeof(T) * n); return; }
    shim_detail::World &w = c.world();
    w.slot[c.rank()] = in;
    w.barrier();
    for (int s = 0; s < c.size(); s++)
        std::memcpy(out + (size_t)s * n, static_cast<const T *>(w.slot[s]) + (size_t)c.rank() * n, sizeof(T) * n);
    w.barrier();
}
template <class T, class Op> void reduce(const communicator &c, const T *in, int n, T *out, Op op, int root) {
    if (c.size() == 1) { if (out) std::memcpy(out, in, sizeof(T) * n); return; }
    shim_detail::World &w = c.world();
    w.slot[c.rank()] = in;
    w.barrier();
    if (c.rank() == root) {
        std::vector<T> acc(static_cast<const T *>(w.slot[0]), static_cast<const T *>(w.slot[0]) + n);
        for (int s = 1; s < c.size(); s++) {
            const T *p = static_cast<const T *>(w.slot[s]);
            for (int i = 0; i < n; i++) acc[i] = op(acc[i], p[i]);
        }
        std::memcpy(out, acc.data(), sizeof(T) * n);  // (out may alias in)
    }
    w.barrier();
}
template <class T, class Op> void reduce(const communicator &c, const T &in, T &out, Op op, int root) { reduce(c, &in, 1, &out, op, root); }
template <class T, class Op> void all_reduce(const communicator &c, const T *in, int n, T *out, Op op) {
    if (c.size() == 1) { std::memcpy(out, in, sizeof(T) * n); return; }
    shim_detail::World &w = c.world();
    w.slot[c.rank()] = in;
    w.barrier();
    std::vector<T> acc(static_cast<const T *>(w.slot[0]), static_cast<const T *>(w.slot[0]) + n);
    for (int s = 1; s < c.size(); s++) {
        const T *p = static_cast<const T *>(w.slot[s]);
        for (int i = 0; i < n; i++) acc[i] = op(acc[i], p[i]);
    }
    w.barrier();
    std::memcpy(out, acc.data(), sizeof(T) * n);
    w.barrier();
}
template <class T, class Op> void all_reduce(const communicator &c, const T &in, T &out, Op op) { all_reduce(c, &in, 1, &out, op); }
template <class T> void broadcast(const communicator &c, T *buf, int n, int root) {
    if (c.size() == 1) return;
    shim_detail::World &w = c.world();
    if (c.rank() == root) w.slot[root] = buf;
    w.barrier();
    if (c.rank() != root) std::memcpy(buf, w.slot[root], sizeof(T) * n);
    w.barrier();
}
template <class T> void broadcast(const communicator &c, T &v, int root) { broadcast(c, &v, 1, root); }
template <class T> void all_gather(const communicator &c, const T &in, std::vector<T> &out) {
    if (c.size() == 1) { out.assign(1, in); return; }
    shim_detail::World &w = c.world();
    w.slot[c.rank()] = &in;
    w.barrier();
    std::vector<T> tmp;
    for (int s = 0; s < c.size(); s++) tmp.push_back(*static_cast<const T *>(w.slot[s]));
    w.barrier();
    out.swap(tmp);
}
template <class T> void gather(const communicator &c, const T &in, std::vector<T> &out, int root) {
    std::vector<T> all;
    all_gather(c, in, all);
    if (c.rank() == root) out.swap(all);
}
}}  // namespace boost::mpi
#endif
