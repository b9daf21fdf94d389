// OUT-OF-PATH stand-in (test infrastructure): graph regularisation is refused at the GPU boundary and never taken
// by the configurations the oracle is checked on. Same signature as the reference's function; calling it throws.
#pragma once
#include <FactorNet/core/types.hpp>
#include <stdexcept>
namespace FactorNet { namespace features {
template<typename Scalar>
inline void apply_graph_reg(DenseMatrix<Scalar>&, const SparseMatrix<Scalar>&, const DenseMatrix<Scalar>&, Scalar) {
    throw std::logic_error("apply_graph_reg: outside the compiled path");
}
}}
