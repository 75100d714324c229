#include "common.cuh"
#include <stdio.h>
static char g_last_msg[512] = "";
extern "C" void mvae_set_last_cuda_error(int code, const char* file, int line) {
  snprintf(g_last_msg, sizeof(g_last_msg), "%s:%d: %s (%d)", file, line, cudaGetErrorString((cudaError_t)code), code);
}
extern "C" const char* mvae_last_cuda_error(void) { return g_last_msg; }
extern "C" const char* mvae_strerror(int rc) {
  switch (rc) {
    case MVAE_OK: return "ok";
    case MVAE_ERR_INVALID: return "invalid argument";
    case MVAE_ERR_WORKSPACE: return "workspace too small";
    case MVAE_ERR_CUDA: return "CUDA runtime error (see mvae_last_cuda_error)";
    case MVAE_ERR_UNSUPPORTED: return "unsupported configuration";
    case MVAE_ERR_DRIVER: return "CUDA driver entry point / tensor-map encode failed";
    default: return "unknown error";
  }
}
