// abi.cu — the parts of the C ABI that are not kernels: error text, version, launch counter.
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "common.cuh"

namespace dbs {
static thread_local char t_error[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}
}  // namespace dbs

namespace dbs {
bool g_prof_on = false;
bool pdl_enabled() {
  // Measured at 64x64, graph-replayed EM steps (profiles/README.md): 0.611 -> 0.583 ms per step at B = 64 and 0.914 -> 0.898 at
  // B = 128, but 1.55 -> 1.58 ms at B = 256 and 11.9 -> 12.2 ms at B = 2048 — CTAs of the next grid parked on the SMs cost the
  // running persistent kernels more than the hidden prologues save once a kernel lasts longer than a few waves.  Off unless
  // DBS_PDL=1 (small-batch serving).
  static const bool on = [] { const char* e = getenv("DBS_PDL"); return e && e[0] == '1'; }();
  return on;
}
namespace {
struct ProfPair { cudaEvent_t a, b; int cls; std::string label; };
std::vector<std::pair<std::string, float>> g_detail;
thread_local char t_label[96] = "";
std::vector<ProfPair> g_pairs;       // recorded pairs awaiting dbs_profile_read
std::vector<cudaEvent_t> g_pool;     // recycled events
std::mutex g_prof_mu;
cudaEvent_t g_open[kProfNumClasses];
double g_ms[kProfNumClasses];
uint64_t g_cnt[kProfNumClasses];
const char* kProfNames[kProfNumClasses] = {"conv_tc", "conv_generic", "gn_stats", "gn_apply", "em_step", "unet_misc",
                                           "ising_sweep", "ising_observables", "ising_misc", "lj_pair",
                                           "csn_linear", "csn_attention"};
cudaEvent_t take_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
}  // namespace
void prof_set_label(const char* fmt, ...) {
  if (!g_prof_on) return;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_label, sizeof(t_label), fmt, ap);
  va_end(ap);
}
void prof_begin(ProfClass c, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_open[c] = take_event();
  cudaEventRecord(g_open[c], st);
}
void prof_end(ProfClass c, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEvent_t b = take_event();
  cudaEventRecord(b, st);
  g_pairs.push_back({g_open[c], b, (int)c, std::string(t_label)});
  t_label[0] = 0;
}
}  // namespace dbs

extern "C" int dbs_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(dbs::g_prof_mu);
  dbs::g_prof_on = on != 0;
  return DBS_OK;
}
extern "C" int dbs_profile_num_classes(void) { return dbs::kProfNumClasses; }
extern "C" const char* dbs_profile_class_name(int c) {
  return (c >= 0 && c < dbs::kProfNumClasses) ? dbs::kProfNames[c] : "";
}
extern "C" int dbs_profile_read(double* ms, uint64_t* launches, int n, int reset) {
  using namespace dbs;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (const ProfPair& pr : g_pairs) {
    float t = 0.f;
    if (cudaEventSynchronize(pr.b) != cudaSuccess || cudaEventElapsedTime(&t, pr.a, pr.b) != cudaSuccess) {
      set_error("profile_read: event query failed");
      return DBS_ERR_CUDA;
    }
    g_ms[pr.cls] += t; g_cnt[pr.cls] += 1;
    if (!pr.label.empty()) g_detail.push_back({std::string(kProfNames[pr.cls]) + " " + pr.label, t});
    g_pool.push_back(pr.a); g_pool.push_back(pr.b);
  }
  g_pairs.clear();
  for (int c = 0; c < n && c < kProfNumClasses; ++c) { if (ms) ms[c] = g_ms[c]; if (launches) launches[c] = g_cnt[c]; }
  if (reset) for (int c = 0; c < kProfNumClasses; ++c) { g_ms[c] = 0; g_cnt[c] = 0; }
  return DBS_OK;
}

// Writes "class label<TAB>ms" lines for every labelled launch seen by dbs_profile_read since the
// last dump, then clears the list.
extern "C" int dbs_profile_dump(const char* path) {
  using namespace dbs;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  FILE* f = fopen(path, "w");
  if (!f) { set_error("profile_dump: cannot open %s", path); return DBS_ERR_INVALID; }
  for (auto& d : g_detail) fprintf(f, "%s\t%.4f\n", d.first.c_str(), d.second);
  fclose(f);
  g_detail.clear();
  return DBS_OK;
}

extern "C" const char* dbs_last_error(void) { return dbs::t_error; }
extern "C" int dbs_abi_version(void) { return 1000; }
extern "C" int dbs_built_sm(void) { return 100; }
extern "C" uint64_t dbs_launch_count(void) { return dbs::g_launches.load(std::memory_order_relaxed); }
