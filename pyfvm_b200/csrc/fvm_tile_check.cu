// Offline compile check of the tile-kernel template (make check): the
// convection-diffusion functor of BASELINE config 4 as codegen.py prints it.
#include "fvm_tile_args.h"
#define FVM_MODE 0
#define FVM_DIM 2
#define FVM_THREADS 256
#define FVM_EDGE_USES_X 1
#define FVM_EDGE_USES_EL 0
#define FVM_EDGE_USES_U 0
#define FVM_EDGE_MASKED 0
#define FVM_VERT_USES_X 0
#define FVM_VERT_USES_U 0
#define FVM_VERT_USES_CV 1
#define FVM_VERT_MASKED 0
#define FVM_HAS_VERTEX 1
#define FVM_HAS_DIRICHLET 1

__device__ __forceinline__ void fvm_edge(const double uk0, const double uk1,
    const double x0_0, const double x0_1, const double x0_2,
    const double x1_0, const double x1_1, const double x1_2,
    const double edge_ce_ratio, const double edge_length, const unsigned mask,
    double& D, double& O, double& C) {
  {
    const double t0 = (1.0/2.0)*edge_ce_ratio;
    const double t1 = edge_ce_ratio*x0_0;
    const double t2 = t0*x0_1;
    D += edge_ce_ratio*x1_0 + edge_ce_ratio + t0*x1_1 - t1 - t2;
    O += edge_ce_ratio*x1_0 + (1.0/2.0)*edge_ce_ratio*x1_1 - edge_ce_ratio - t1 - t2;
  }
}

__device__ __forceinline__ void fvm_vertex(const double uk0, const double control_volume,
    const double x_0, const double x_1, const double x_2, const unsigned mask,
    double& L, double& C) {
  {
    C += -1.0*control_volume;
  }
}

__device__ __forceinline__ void fvm_dirichlet(const int id, const double uk0,
    const double x_0, const double x_1, const double x_2,
    double& coeff, double& val) {
  switch (id) {
  case 1:
    {
      coeff += 1;
    }
    break;
  default: break;
  }
}
#include "fvm_tile_kernel.cuh"
