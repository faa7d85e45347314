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

// solver/uris.cpp cannot be compiled here (it pulls in the VTK readers), but fluid.cpp / fsi.cpp call one routine of it per element:
// uris::eval_uris_ris_factors_quadrature (uris.cpp:1577-1673).  RESTATEMENT of that routine (test infrastructure; parity of the
// URIS factor itself is therefore against this restatement, while the URIS terms of fluid_3d_m / fluid_3d_c that consume it are the
// compiled reference's).  It reads the same urisType members the reference reads; the harness fills them (svref_set_uris).
#include "uris.h"
#include <cmath>
namespace uris {
void eval_uris_ris_factors_quadrature(const ComMod& cm, const mshType& lM, const fsType& fs, const int e,
    Vector<double>& factor, Array<double>& velTerm)
{
  const int nU = cm.nUris, nsd = cm.nsd;
  factor.resize(fs.nG); factor = 0.0;                                   // :1587-1590
  velTerm.resize(nsd, fs.nG); velTerm = 0.0;
  if (!cm.urisActFlag) return;                                          // :1592-1594
  const double pi = 3.141592653589793238462643383279502884;
  for (int g = 0; g < fs.nG; g++) {
    for (int iU = 0; iU < nU; iU++) {
      const auto& u = cm.uris[iU];
      double dist = 0.0, dsc = 0.0, vel[3] = {0.0, 0.0, 0.0};
      for (int a = 0; a < fs.eNoN; a++) {                               // :1604-1616
        const int Ac = lM.IEN(a, e);
        dist += fs.N(a, g) * std::fabs(u.sdf(Ac));
        if (u.scaffold_flag) dsc += fs.N(a, g) * std::fabs(u.scaffold_udf(Ac));
        if (u.include_uris_velocity)
          for (int i = 0; i < nsd; i++) vel[i] = vel[i] + fs.N(a, g) * u.valve_velocity_fluid(i, Ac);
      }
      // half-thickness: ramp between the open and the closed value over the DxOpen / DxClose steps (:1625-1649)
      const double d0 = u.clsFlg ? u.sdf_deps : u.sdf_deps_close, d1 = u.clsFlg ? u.sdf_deps_close : u.sdf_deps;
      const int nSteps = u.clsFlg ? u.DxClose.nslices() : u.DxOpen.nslices();
      double deps;
      if (nSteps <= 0 || u.cnt >= nSteps) deps = d1;
      else if (u.cnt <= 0) deps = d0;
      else deps = d0 + (static_cast<double>(u.cnt) / static_cast<double>(nSteps)) * (d1 - d0);
      double delta = 0.0, deltaSc = 0.0;
      if (dist < deps && deps > 0.0) delta = (1 + std::cos(pi * dist / deps)) / (2 * deps * deps);          // :1650-1653
      if (u.scaffold_flag) {                                                                                // :1654-1663
        const double sd = u.sdf_deps_close;
        if (dsc < sd && sd > 0.0) deltaSc = (1 + std::cos(pi * dsc / sd)) / (2 * sd * sd);
      }
      factor(g) += u.resistance * (delta + deltaSc);                                                        // :1665
      if (u.include_uris_velocity)
        for (int i = 0; i < nsd; i++) velTerm(i, g) = velTerm(i, g) + u.resistance * delta * vel[i];        // :1667-1670
    }
  }
}
}
