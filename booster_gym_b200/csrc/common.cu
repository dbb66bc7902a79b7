// common.cu - last-error string, version and ABI self-check of libb200t1.so
#include <string.h>

#include "common.cuh"

namespace b200 {
static thread_local char g_err[512] = "";
int set_error(int code, const char* msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}
int set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof g_err, "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), where);
    return B200_ERR_CUDA;
}
}  // namespace b200

extern "C" {
const char* b200_last_error(void) { return b200::g_err; }
int b200_version(void) { return 100; }
int b200_sizeof(int what) {
    switch (what) {
        case 0: return (int)sizeof(B200T1ModelF);
        case 1: return (int)sizeof(B200T1Config);
        case 2: return (int)sizeof(B200PpoConfig);
        case 3: return (int)sizeof(B200T1ModelD);
    }
    return -1;
}
}
