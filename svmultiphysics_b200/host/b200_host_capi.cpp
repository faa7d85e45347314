// Flat entry points of the C++ host layer, for drivers that hold the reference's objects behind opaque pointers
// (the parity harness loads this library next to the compiled reference and runs the reference's own call sequence
// ls_alloc -> global_eq_assem -> ls_solve through B200LinearAlgebra instead of FsilsLinearAlgebra).
#include "B200LinearAlgebra.h"

extern "C" {

__attribute__((visibility("default")))
void* b200host_new(int device, int scatter)
{
  auto* la = new B200LinearAlgebra();
  la->device = device;
  la->scatter = scatter;
  return static_cast<LinearAlgebra*>(la);
}

__attribute__((visibility("default")))
void b200host_delete(void* p)
{
  delete dynamic_cast<B200LinearAlgebra*>(static_cast<LinearAlgebra*>(p));
}

/// The early-out of eq_assem::global_eq_assem.  Returns 1 when the mesh was assembled on the device.
__attribute__((visibility("default")))
int b200host_global_eq_assem(void* com_mod, void* cep_mod, const void* msh, const void* solutions)
{
  return b200::global_eq_assem(*static_cast<ComMod*>(com_mod), *static_cast<CepMod*>(cep_mod),
                               *static_cast<const mshType*>(msh), *static_cast<const SolutionStates*>(solutions)) ? 1 : 0;
}

__attribute__((visibility("default")))
void b200host_download(void* p, int what, double* dst)
{
  auto* la = dynamic_cast<B200LinearAlgebra*>(static_cast<LinearAlgebra*>(p));
  if (!la) throw std::runtime_error("b200host_download: not a B200LinearAlgebra");
  la->download(what, dst);
}

/// Where Integrator::step calls ustruct::ustruct_r (Integrator.cpp:135-137).
__attribute__((visibility("default")))
void b200host_ustruct_r(void* p, void* com_mod)
{
  auto* la = dynamic_cast<B200LinearAlgebra*>(static_cast<LinearAlgebra*>(p));
  if (!la) throw std::runtime_error("b200host_ustruct_r: not a B200LinearAlgebra");
  la->ustruct_r(*static_cast<ComMod*>(com_mod));
}

/// Where Integrator::step calls fs::thood_val_rc (Integrator.cpp:140-147).
__attribute__((visibility("default")))
void b200host_thood_val_rc(void* p)
{
  auto* la = dynamic_cast<B200LinearAlgebra*>(static_cast<LinearAlgebra*>(p));
  if (!la) throw std::runtime_error("b200host_thood_val_rc: not a B200LinearAlgebra");
  la->thood_val_rc();
}

__attribute__((visibility("default")))
long long b200host_launch_count(void* p)
{
  auto* la = dynamic_cast<B200LinearAlgebra*>(static_cast<LinearAlgebra*>(p));
  return (la && la->ctx) ? (long long)svb200_launch_count(la->ctx) : -1;
}

}  // extern "C"
