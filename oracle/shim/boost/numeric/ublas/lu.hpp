/* boost/numeric/ublas/lu.hpp — minimal SHIM for the reference's src/sample/center_of_mass.cpp (determinant of a 3x3 product):
 * permutation_matrix<T>(n) (identity on construction, (i) access, size()) and lu_factorize(m, pm) as uBLAS does it -- in place,
 * column by column, pivot = FIRST entry of largest magnitude at or below the diagonal, rows swapped and recorded in pm(i),
 * multipliers stored below the diagonal, trailing block updated; returns 0, or i + 1 for the first zero pivot column. */
#ifndef ORACLE_SHIM_UBLAS_LU_HPP
#define ORACLE_SHIM_UBLAS_LU_HPP
#include <cmath>
#include <vector>

#include <boost/numeric/ublas/matrix.hpp>
namespace boost { namespace numeric { namespace ublas {
template <class T>
class permutation_matrix {
    std::vector<T> p_;
   public:
    explicit permutation_matrix(std::size_t n) : p_(n) {
        for (std::size_t i = 0; i < n; i++) p_[i] = (T)i;
    }
    std::size_t size() const { return p_.size(); }
    T &operator()(std::size_t i) { return p_[i]; }
    const T &operator()(std::size_t i) const { return p_[i]; }
};
template <class M, class P>
std::size_t lu_factorize(M &m, P &pm) {
    std::size_t singular = 0;
    const std::size_t size1 = m.size1(), size2 = m.size2(), size = size1 < size2 ? size1 : size2;
    for (std::size_t i = 0; i < size; ++i) {
        std::size_t ip = i;
        double best = std::fabs(m(i, i));
        for (std::size_t k = i + 1; k < size1; k++)
            if (std::fabs(m(k, i)) > best) {
                best = std::fabs(m(k, i));
                ip = k;
            }
        if (m(ip, i) != 0.0) {
            if (ip != i) {
                pm(i) = ip;
                for (std::size_t j = 0; j < size2; j++) {
                    const double t = m(i, j);
                    m(i, j) = m(ip, j);
                    m(ip, j) = t;
                }
            }
            const double inv = 1.0 / m(i, i);
            for (std::size_t k = i + 1; k < size1; k++) m(k, i) *= inv;
        } else if (singular == 0) {
            singular = i + 1;
        }
        for (std::size_t k = i + 1; k < size1; k++)
            for (std::size_t j = i + 1; j < size2; j++) m(k, j) -= m(k, i) * m(i, j);
    }
    return singular;
}
}}}  // namespace boost::numeric::ublas
#endif
