// Library-wide entry points: version, thread-local error string, device query.
#include "common.cuh"
#include "../../include/meshrcnn_b200.h"
#include <stdarg.h>
#include <string.h>

namespace mrb {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace mrb

extern "C" int mrb_version(void) { return MRB_VERSION; }

extern "C" const char* mrb_last_error(void) { return mrb::g_err; }

extern "C" int mrb_device_info(int* sm_major, int* sm_minor, int* num_sms) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        mrb::set_error("cudaGetDevice: %s", cudaGetErrorString(e));
        return MRB_ERR_CUDA;
    }
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) {
        mrb::set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
        return MRB_ERR_CUDA;
    }
    if (sm_major) *sm_major = p.major;
    if (sm_minor) *sm_minor = p.minor;
    if (num_sms) *num_sms = p.multiProcessorCount;
    return MRB_OK;
}
