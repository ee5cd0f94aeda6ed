/* boost/thread.hpp — SHIM of the Boost.Thread pieces the reference's scatter devices use (worker threads fed by a queue and
 * a barrier, stopped with interrupt()), over <thread>.  Interruption is a flag checked at the waits, as in Boost. */
#ifndef ORACLE_SHIM_BOOST_THREAD_HPP
#define ORACLE_SHIM_BOOST_THREAD_HPP
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
namespace boost {
struct thread_interrupted {};
namespace shim_detail {
struct tstate {
    std::atomic<bool> interrupt{false};
    std::mutex m;                                 // guards `waiting`
    std::condition_variable *waiting = nullptr;   // the condition variable this thread is blocked on, if any
    void request() {
        interrupt = true;
        std::lock_guard<std::mutex> l(m);
        if (waiting) waiting->notify_all();
    }
};
inline tstate *&current() {
    static thread_local tstate *p = nullptr;
    return p;
}
inline void interruption_point() {
    if (current() && current()->interrupt.load()) throw thread_interrupted();
}
// blocks on cv (woken by notify; a 20 ms timeout only covers the window in which an interrupt request could be missed)
inline void interruptible_wait(std::condition_variable &cv, std::unique_lock<std::mutex> &l) {
    interruption_point();
    tstate *st = current();
    if (st) {
        std::lock_guard<std::mutex> g(st->m);
        st->waiting = &cv;
    }
    cv.wait_for(l, std::chrono::milliseconds(20));
    if (st) {
        std::lock_guard<std::mutex> g(st->m);
        st->waiting = nullptr;
    }
    interruption_point();
}
}  // namespace shim_detail
using std::bind;
template <class T> using shared_ptr = std::shared_ptr<T>;
class condition_variable;
class mutex {
    std::mutex m_;
   public:
    class scoped_lock {
        std::unique_lock<std::mutex> l_;
        friend class condition_variable;
       public:
        explicit scoped_lock(mutex &m) : l_(m.m_) {}
        void unlock() { l_.unlock(); }
        void lock() { l_.lock(); }
    };
};
class condition_variable {
    std::condition_variable cv_;
   public:
    void notify_all() { cv_.notify_all(); }
    void notify_one() { cv_.notify_one(); }
    // an interruption point; may wake spuriously (callers loop on their predicate)
    void wait(mutex::scoped_lock &l) { shim_detail::interruptible_wait(cv_, l.l_); }
};
class barrier {
    std::mutex m_;
    std::condition_variable cv_;
    unsigned threshold_, count_, generation_;
   public:
    explicit barrier(unsigned n) : threshold_(n), count_(n), generation_(0) {}
    bool wait() {
        std::unique_lock<std::mutex> l(m_);
        const unsigned gen = generation_;
        if (--count_ == 0) {
            generation_++;
            count_ = threshold_;
            cv_.notify_all();
            return true;
        }
        while (gen == generation_) shim_detail::interruptible_wait(cv_, l);
        return false;
    }
};
class thread {
    std::shared_ptr<shim_detail::tstate> st_;
    std::thread t_;
   public:
    typedef std::thread::id id;
    static unsigned hardware_concurrency() { return std::thread::hardware_concurrency(); }
    template <class F>
    explicit thread(F f) : st_(new shim_detail::tstate) {
        std::shared_ptr<shim_detail::tstate> st = st_;
        t_ = std::thread([st, f]() mutable {
            shim_detail::current() = st.get();
            try {
                f();
            } catch (thread_interrupted &) {
            }
        });
    }
    ~thread() {  // (Boost detaches; here the interrupted worker is joined so that nothing outlives the device)
        if (t_.joinable()) {
            st_->request();
            t_.join();
        }
    }
    void interrupt() { st_->request(); }
    void join() { if (t_.joinable()) t_.join(); }
    id get_id() const { return t_.get_id(); }
};
namespace posix_time {
struct milliseconds {
    long n;
    explicit milliseconds(long v) : n(v) {}
};
}  // namespace posix_time
namespace this_thread {
inline std::thread::id get_id() { return std::this_thread::get_id(); }
inline void sleep(const posix_time::milliseconds &d) { std::this_thread::sleep_for(std::chrono::milliseconds(d.n)); }
inline void interruption_point() { shim_detail::interruption_point(); }
}  // namespace this_thread
}  // namespace boost
#endif
