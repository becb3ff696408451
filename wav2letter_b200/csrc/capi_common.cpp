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
// event LIST mode: the k-th profiled launch of the selected kind records pair k
static thread_local cudaEvent_t* g_list_start = nullptr;
static thread_local cudaEvent_t* g_list_stop = nullptr;
static thread_local int g_list_n = 0, g_list_used = 0, g_list_kind = 0, g_cur_kind = 0;
static thread_local bool g_list_armed = false;
void profile_kind(int kind) { g_cur_kind = kind; }
void profile_start(cudaStream_t s) {
  if (g_list_start) {
    g_list_armed = (g_list_kind == 0 || g_list_kind == g_cur_kind) && g_list_used < g_list_n;
    if (g_list_armed) cudaEventRecord(g_list_start[g_list_used], s);
  } else if (g_ev_start) {
    cudaEventRecord(g_ev_start, s);
  }
}
void profile_stop(cudaStream_t s) {
  if (g_list_start) {
    if (g_list_armed) cudaEventRecord(g_list_stop[g_list_used++], s);
    g_list_armed = false;
  } else if (g_ev_stop) {
    cudaEventRecord(g_ev_stop, s);
  }
  g_cur_kind = 0;
}
void set_profile_events(cudaEvent_t a, cudaEvent_t b) {
  g_ev_start = a;
  g_ev_stop = b;
  g_list_start = g_list_stop = nullptr;
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
int w2l_set_profile_event_list(int kind, void** starts, void** stops, int n) {
  w2l::g_list_start = reinterpret_cast<cudaEvent_t*>(starts);
  w2l::g_list_stop = reinterpret_cast<cudaEvent_t*>(stops);
  w2l::g_list_n = starts ? n : 0;
  w2l::g_list_used = 0;
  w2l::g_list_kind = kind;
  w2l::g_ev_start = w2l::g_ev_stop = nullptr;
  return 0;
}
int w2l_profile_events_used(void) { return w2l::g_list_used; }
}
