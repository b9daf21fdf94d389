// OUT-OF-PATH stand-in (test infrastructure): SVD-based initialisation is outside the hot path;
// nmf_init.hpp only needs the entry point to exist. Calling it throws.
#pragma once
#include <FactorNet/core/svd_config.hpp>
#include <FactorNet/core/svd_result.hpp>
#include <FactorNet/core/types.hpp>
#include <stdexcept>
namespace FactorNet { namespace svd {
template <typename MatrixType, typename Scalar>
SVDResult<Scalar> lanczos_svd(const MatrixType&, const SVDConfig<Scalar>&) {
    throw std::logic_error("lanczos_svd: SVD initialisation is outside the compiled path");
}
}}
