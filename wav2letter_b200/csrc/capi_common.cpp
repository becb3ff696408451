// capi_common.cpp — error text, launch counter, version.
#include <string>

#include "common.cuh"

namespace w2l {
static thread_local std::string g_err;
static thread_local long long g_launches = 0;

void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
void count_launch(int n) { g_launches += n; }
}  // namespace w2l

extern "C" {
int w2l_version(void) { return 100; }
const char* w2l_last_error(void) { return w2l::g_err.c_str(); }
long long w2l_launch_count(void) { return w2l::g_launches; }
void w2l_reset_launch_count(void) { w2l::g_launches = 0; }
}
