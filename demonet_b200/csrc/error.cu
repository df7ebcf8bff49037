#include "common.cuh"

namespace dn {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace dn

extern "C" const char* dn_last_error(void) { return dn::g_err; }
extern "C" int dn_abi_version(void) { return DN_ABI_VERSION; }
