// OUT-OF-PATH stand-in (test infrastructure): IRLS losses (non-MSE) are outside the hot path. Same signatures as the
// reference's functions; calling them throws.
#pragma once
#include <FactorNet/core/types.hpp>
#include <FactorNet/core/config.hpp>
#include <stdexcept>
namespace FactorNet { namespace primitives {
namespace detail {
template<typename Scalar>
inline Scalar compute_irls_weight(Scalar, Scalar, const LossConfig<Scalar>&, Scalar = static_cast<Scalar>(0), Scalar = static_cast<Scalar>(0)) {
    throw std::logic_error("compute_irls_weight: outside the compiled path");
}
}
template<typename Scalar, typename SparseMatType>
void nnls_batch_irls_sparse(const SparseMatType&, const DenseMatrix<Scalar>&, const DenseMatrix<Scalar>&, DenseMatrix<Scalar>&,
                            const LossConfig<Scalar>&, Scalar, Scalar, bool, int, Scalar, int, Scalar, int,
                            const Scalar* = nullptr, const Scalar* = nullptr) {
    throw std::logic_error("nnls_batch_irls_sparse: outside the compiled path");
}
template<typename Scalar, typename DenseMatType>
void nnls_batch_irls_dense(const DenseMatType&, const DenseMatrix<Scalar>&, const DenseMatrix<Scalar>&, DenseMatrix<Scalar>&,
                           const LossConfig<Scalar>&, Scalar, Scalar, bool, int, Scalar, int, Scalar, int,
                           const Scalar* = nullptr, const Scalar* = nullptr) {
    throw std::logic_error("nnls_batch_irls_dense: outside the compiled path");
}
}}
