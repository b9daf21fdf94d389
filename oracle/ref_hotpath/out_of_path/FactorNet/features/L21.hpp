// OUT-OF-PATH stand-in (test infrastructure): L21 is outside the hot path. Calling it throws.
#pragma once
#include <FactorNet/core/types.hpp>
#include <stdexcept>
namespace FactorNet { namespace features {
template<typename Scalar>
inline void apply_L21(DenseMatrix<Scalar>&, const DenseMatrix<Scalar>&, Scalar) {
    throw std::logic_error("apply_L21: outside the compiled path");
}
}}
