/* metis_shim.c — TEST INFRASTRUCTURE ONLY: one flat entry point around the serial METIS that the reference vendors
 * (Code/ThirdParty/metis_internal, GKlib in Code/ThirdParty/gklib_internal), compiled where those sources lie by
 * `make -C oracle metis` into oracle/_ref/libsvmetis.so.
 *
 * The reference partitions with ParMETIS_V3_PartMeshKway (Code/Source/solver/SPLIT.c:87: dual graph with
 * ncommonnodes = eNoNb, seed 10, imbalance 1.05), which needs one MPI process per part; MPI is absent here, so the
 * multi-GPU parity tests take the same kind of partition — k-way on the dual graph, same ncommon, seed and imbalance —
 * from METIS_PartMeshDual of the reference's own METIS and inject it as the part[] array that part_msh otherwise
 * computes (Code/Source/solver/distribute.cpp:2209), as SURVEY.md 8(c) proposes. */
#include <stddef.h>
#include <stdlib.h>
#include "metis.h"

int svmetis_part_mesh_dual(int ne, int nn, int eNoN, const int* IEN, int ncommon, int nparts, int seed, int* epart, int* npart)
{
  idx_t options[METIS_NOPTIONS];
  idx_t ne_ = ne, nn_ = nn, nc = ncommon, np_ = nparts, objval = 0;
  idx_t* eptr = (idx_t*)malloc(sizeof(idx_t) * ((size_t)ne + 1));
  if (!eptr) return -1;
  for (int e = 0; e <= ne; e++) eptr[e] = (idx_t)e * eNoN;
  METIS_SetDefaultOptions(options);
  options[METIS_OPTION_SEED] = seed;
  options[METIS_OPTION_NUMBERING] = 0;
  options[METIS_OPTION_UFACTOR] = 50;       /* 1.05, UNBALANCE_FRACTION of ParMETIS */
  int rc = METIS_PartMeshDual(&ne_, &nn_, eptr, (idx_t*)IEN, NULL, NULL, &nc, &np_, NULL, options, &objval, (idx_t*)epart, (idx_t*)npart);
  free(eptr);
  return rc == METIS_OK ? (int)objval : -2;
}
