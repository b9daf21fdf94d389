// OUT-OF-PATH stand-in (test infrastructure): the angular penalty is outside the hot path. Calling it throws.
#pragma once
#include <FactorNet/core/types.hpp>
#include <stdexcept>
namespace FactorNet { namespace features {
template<typename Scalar>
inline void apply_angular(DenseMatrix<Scalar>&, const DenseMatrix<Scalar>&, Scalar) {
    throw std::logic_error("apply_angular: outside the compiled path");
}
template<typename Scalar>
inline void apply_angular_posthoc(DenseMatrix<Scalar>&, Scalar) {
    throw std::logic_error("apply_angular_posthoc: outside the compiled path");
}
}}
