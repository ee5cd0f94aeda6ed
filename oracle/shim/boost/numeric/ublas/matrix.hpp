/* boost/numeric/ublas/matrix.hpp — minimal SHIM of the uBLAS dense matrix for the reference's src/sample/motion_walker.cpp:
 * matrix<T>(r, c) with (i, j) access, identity_matrix<T>, prod(matrix, matrix) (inner sum in index order, as uBLAS does).
 * decomposition_plan.hpp includes the header without using it.  For src/sample/center_of_mass.cpp (the rotational fit):
 * zero_matrix, outer_prod, scalar * matrix, matrix += matrix, matrix x column vector, vector[i], raw row-major storage. */
#ifndef ORACLE_SHIM_UBLAS_MATRIX_HPP
#define ORACLE_SHIM_UBLAS_MATRIX_HPP
#include <cstddef>
#include <vector>
namespace boost { namespace numeric { namespace ublas {
namespace detail {}  /* (all_vectors_scatter_device.cpp has a using-directive for it) */
template <class T>
class identity_matrix {
    std::size_t n_;
   public:
    explicit identity_matrix(std::size_t n) : n_(n) {}
    identity_matrix(std::size_t n, std::size_t) : n_(n) {}
    std::size_t size1() const { return n_; }
};
template <class T>
class matrix {
    std::size_t r_, c_;
    std::vector<T> d_;
   public:
    matrix() : r_(0), c_(0) {}
    matrix(std::size_t r, std::size_t c) : r_(r), c_(c), d_(r * c) {}
    matrix(const identity_matrix<T> &I) { *this = I; }
    T *data() { return d_.data(); }  /* row-major storage */
    matrix &operator+=(const matrix &o) {
        for (std::size_t i = 0; i < d_.size(); i++) d_[i] += o.d_[i];
        return *this;
    }
    matrix &operator=(const identity_matrix<T> &I) {
        r_ = c_ = I.size1();
        d_.assign(r_ * c_, T(0));
        for (std::size_t i = 0; i < r_; i++) d_[i * c_ + i] = T(1);
        return *this;
    }
    T &operator()(std::size_t i, std::size_t j) { return d_[i * c_ + j]; }
    const T &operator()(std::size_t i, std::size_t j) const { return d_[i * c_ + j]; }
    std::size_t size1() const { return r_; }
    std::size_t size2() const { return c_; }
};
template <class T>
matrix<T> prod(const matrix<T> &a, const matrix<T> &b) {
    matrix<T> c(a.size1(), b.size2());
    for (std::size_t i = 0; i < a.size1(); i++)
        for (std::size_t j = 0; j < b.size2(); j++) {
            T t = T(0);
            for (std::size_t k = 0; k < a.size2(); k++) t += a(i, k) * b(k, j);
            c(i, j) = t;
        }
    return c;
}
/* dense vector and row-vector x matrix product (coordinate_set.cpp:161-173), inner sum in index order */
template <class T>
class vector {
    std::vector<T> d_;
   public:
    vector() {}
    explicit vector(std::size_t n) : d_(n) {}
    T &operator()(std::size_t i) { return d_[i]; }
    const T &operator()(std::size_t i) const { return d_[i]; }
    T &operator[](std::size_t i) { return d_[i]; }
    const T &operator[](std::size_t i) const { return d_[i]; }
    std::size_t size() const { return d_.size(); }
    T *data() { return d_.data(); }
};
template <class T>
matrix<T> zero_matrix(std::size_t r, std::size_t c) {
    return matrix<T>(r, c);  /* value-initialised */
}
template <class T>
matrix<T> outer_prod(const vector<T> &u, const vector<T> &v) {
    matrix<T> m(u.size(), v.size());
    for (std::size_t i = 0; i < u.size(); i++)
        for (std::size_t j = 0; j < v.size(); j++) m(i, j) = u(i) * v(j);
    return m;
}
template <class T>
matrix<T> operator*(T f, const matrix<T> &a) {
    matrix<T> m(a.size1(), a.size2());
    for (std::size_t i = 0; i < a.size1(); i++)
        for (std::size_t j = 0; j < a.size2(); j++) m(i, j) = f * a(i, j);
    return m;
}
template <class T>
vector<T> prod(const matrix<T> &m, const vector<T> &v) {
    vector<T> r(m.size1());
    for (std::size_t i = 0; i < m.size1(); i++) {
        T t = T(0);
        for (std::size_t j = 0; j < m.size2(); j++) t += m(i, j) * v(j);
        r(i) = t;
    }
    return r;
}
template <class T>
vector<T> prod(const vector<T> &v, const matrix<T> &m) {
    vector<T> r(m.size2());
    for (std::size_t j = 0; j < m.size2(); j++) {
        T t = T(0);
        for (std::size_t i = 0; i < m.size1(); i++) t += v(i) * m(i, j);
        r(j) = t;
    }
    return r;
}
}}}  // namespace boost::numeric::ublas
#endif
