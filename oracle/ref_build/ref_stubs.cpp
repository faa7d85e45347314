// Link-closure helpers for oracle/_ref/libsvref.so (test infrastructure only).
//
// The reference's LinearAlgebra.cpp also instantiates the PETSc and Trilinos adapters, which need
// external libraries that are not in this image.  The oracle only ever uses the FSILS backend, so
// this file supplies the few LinearAlgebra base-class members that FsilsLinearAlgebra needs
// (declared in solver/LinearAlgebra.h:13-37) instead of compiling LinearAlgebra.cpp.
#include "LinearAlgebra.h"
#include <stdexcept>

const std::map<std::string, consts::LinearAlgebraType> LinearAlgebra::name_to_type = {
  {"none", consts::LinearAlgebraType::none},
  {"fsils", consts::LinearAlgebraType::fsils},
};
const std::map<consts::LinearAlgebraType, std::string> LinearAlgebra::type_to_name = {
  {consts::LinearAlgebraType::none, "none"},
  {consts::LinearAlgebraType::fsils, "fsils"},
  {consts::LinearAlgebraType::petsc, "petsc"},
  {consts::LinearAlgebraType::trilinos, "trilinos"},
};
void LinearAlgebra::check_equation_compatibility(const consts::EquationType, const consts::LinearAlgebraType,
    const consts::LinearAlgebraType) {}
LinearAlgebra::LinearAlgebra() {}

// RIS / unfitted-RIS valve models are outside the hot-path scope; the element
// loops only reach them when com_mod.risFlag / urisFlag are set, which the harness never does.
#include "ris.h"
#include "uris.h"
namespace uris {
void eval_uris_ris_factors_quadrature(const ComMod&, const mshType&, const fsType&, const int,
    Vector<double>&, Array<double>&)
{ throw std::runtime_error("[oracle] uris is out of scope"); }
}
namespace ris {
void doassem_ris(ComMod&, const int, const Vector<int>&, const Array3<double>&, const Array<double>&)
{ throw std::runtime_error("[oracle] ris is out of scope"); }
}
