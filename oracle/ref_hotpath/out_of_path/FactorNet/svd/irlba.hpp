// OUT-OF-PATH stand-in (test infrastructure): see svd/lanczos.hpp in this directory.
#pragma once
#include <FactorNet/core/svd_config.hpp>
#include <FactorNet/core/svd_result.hpp>
#include <FactorNet/core/types.hpp>
#include <stdexcept>
namespace FactorNet { namespace svd {
template <typename MatrixType, typename Scalar>
SVDResult<Scalar> irlba_svd(const MatrixType&, const SVDConfig<Scalar>&) {
    throw std::logic_error("irlba_svd: SVD initialisation is outside the compiled path");
}
}}
