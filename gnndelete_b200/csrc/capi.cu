// Error reporting and version of the C ABI.
#include <cstdlib>

#include "common.cuh"

namespace gd {
static thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
}  // namespace gd

extern "C" int gd_version(void) { return 100; /* 0.1.0 */ }
extern "C" const char* gd_last_error(void) { return gd::g_last_error.c_str(); }
extern "C" long long gd_launch_count(void) { return gd::g_launches.load(); }
