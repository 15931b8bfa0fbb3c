// ABI bookkeeping for libfortuna_b200.so.
#include "common.cuh"

extern "C" int fb200_abi_version(void) { return FB200_ABI_VERSION; }

extern "C" const char* fb200_error_string(int code) {
  switch (code) {
    case 0:
      return "ok";
    case FB200_E_NULL:
      return "fortuna_b200: a required pointer is NULL";
    case FB200_E_SHAPE:
      return "fortuna_b200: a size is non-positive or inconsistent";
    case FB200_E_ALIGN:
      return "fortuna_b200: a base pointer is not 16-byte aligned";
    case FB200_E_UNSUPPORTED:
      return "fortuna_b200: shape or dtype outside what the kernels cover";
    case FB200_E_WORKSPACE:
      return "fortuna_b200: workspace too small";
    default:
      break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "fortuna_b200: unknown error code";
}
