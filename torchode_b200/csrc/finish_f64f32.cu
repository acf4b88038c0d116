// explicit instantiation of the finish launcher for data=double, time=float
#include "finish_impl.cuh"
namespace tode {
template int launch_finish<double, float>(const tode_tableau*, const tode_controller*, const tode_state*,
                                     const void* const*, const void*, cudaStream_t);
}
