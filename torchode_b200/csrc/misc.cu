// C-ABI: version, error strings, scratch sizing.
#include <cuda_runtime.h>

#include "../../include/torchode_b200.h"

extern "C" int tode_abi_version(void) { return TODE_ABI_VERSION; }

extern "C" const char* tode_error_string(int code) {
  switch (code) {
    case 0: return "success";
    case TODE_EINVAL: return "invalid argument (NULL pointer, bad size or enum)";
    case TODE_ENOSUP: return "combination not supported by this build";
    case TODE_EALIGN: return "operand not aligned as the layout contract demands (16 bytes)";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "unknown error";
}

// scratch (data dtype elements): 2*B for the initial step (dt0, d1)
extern "C" int64_t tode_scratch_elems(int64_t B, int64_t F) {
  (void)F;
  return 2 * B + 16;
}
