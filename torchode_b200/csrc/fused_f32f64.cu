// explicit instantiation of the fused-solve launcher for data=float, time=double
#include "fused_impl.cuh"
namespace tode {
template int launch_fused<float, double>(int, const double*, const tode_tableau*, const tode_controller*,
                                   const tode_problem*, const tode_solution*, int64_t, cudaStream_t);
}
