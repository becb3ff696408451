// capi_common.cpp — error text, launch counter, version.
#include <string>

#include "common.cuh"

namespace w2l {
void set_profile_events(cudaEvent_t a, cudaEvent_t b);
static thread_local std::string g_err;
static thread_local long long g_launches = 0;

void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
void count_launch(int n) { g_launches += n; }
static thread_local cudaEvent_t g_ev_start = nullptr, g_ev_stop = nullptr;
void profile_start(cudaStream_t s) {
  if (g_ev_start) cudaEventRecord(g_ev_start, s);
}
void profile_stop(cudaStream_t s) {
  if (g_ev_stop) cudaEventRecord(g_ev_stop, s);
}
void set_profile_events(cudaEvent_t a, cudaEvent_t b) {
  g_ev_start = a;
  g_ev_stop = b;
}
}  // namespace w2l

extern "C" {
int w2l_version(void) { return 100; }
const char* w2l_last_error(void) { return w2l::g_err.c_str(); }
long long w2l_launch_count(void) { return w2l::g_launches; }
void w2l_reset_launch_count(void) { w2l::g_launches = 0; }
void w2l_set_profile_events(void* a, void* b) {
  w2l::set_profile_events(static_cast<cudaEvent_t>(a), static_cast<cudaEvent_t>(b));
}
}
