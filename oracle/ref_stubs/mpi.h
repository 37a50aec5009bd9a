/* Stub standing in for OpenMPI's <mpi.h> when compiling the reference's kernel
 * translation units (kernels.cu, kLoss.cu, kDelta.cu, kActivation.cu) as a
 * GPU-side oracle.  The kernel files only need the three MPI datatype
 * constants that GpuTypes.h names in typedef tables (E/GpuTypes.h:84-108);
 * no MPI function is ever called on this path.  Test infrastructure only. */
#ifndef DSB200_STUB_MPI_H
#define DSB200_STUB_MPI_H
typedef int MPI_Datatype;
#define MPI_DOUBLE_PRECISION 1
#define MPI_FLOAT            2
#define MPI_LONG_LONG_INT    3
#endif
