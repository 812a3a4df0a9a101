// Error channel + version for the C ABI (include/dmvae_b200.h).
#include "common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

int dmvae_set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

DMVAE_API const char* dmvae_last_error(void) { return g_err; }
DMVAE_API int dmvae_abi_version(void) { return 8; }

// Device the library was built for; lets the host side fail loudly on anything but sm_100.
DMVAE_API int dmvae_check_device(void) {
    int dev = 0;
    cudaDeviceProp p;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess)
        return dmvae_set_error(DMVAE_ECUDA, "check_device: no CUDA device");
    if (p.major != 10)
        return dmvae_set_error(DMVAE_EUNSUPPORTED, "check_device: built for sm_100a, found sm_%d%d", p.major, p.minor);
    return DMVAE_OK;
}
