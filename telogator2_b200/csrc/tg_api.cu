// Error reporting and device queries of the C ABI (include/telogator2_b200.h).
#include "tg_common.cuh"

static thread_local char g_err[512] = "";

char *tg_err_buf() { return g_err; }

int tg_set_err(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

TG_EXPORT const char *tg_last_error(void) { return g_err; }
TG_EXPORT int tg_version(void) { return 1; }

TG_EXPORT int tg_device_sm_count(int *sm_count)
{
    const int n = tg_sm_count();
    if (n <= 0) return tg_set_err(TG_ECUDA, "no CUDA device available");
    if (sm_count) *sm_count = n;
    return TG_OK;
}
