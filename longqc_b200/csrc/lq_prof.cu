/* lq_prof.cu -- see lq_prof.h */
#include <vector>
#include <string>
#include <string.h>
#include <stdio.h>
#include "lq_prof.h"
#include "lqcov.h"

namespace {
struct Pending { const char *name; cudaEvent_t e0, e1; uint64_t launches, bytes; };
bool g_on = false;
std::vector<Pending> g_pending;
std::vector<LqKernelStat> g_stats;
std::vector<cudaEvent_t> g_pool;
const char *g_open = 0; cudaEvent_t g_e0;
uint64_t g_launches = 0, g_h2d = 0, g_d2h = 0;

cudaEvent_t get_event() { if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; } cudaEvent_t e; cudaEventCreate(&e); return e; }
LqKernelStat *slot(const char *name)
{
    for (size_t i = 0; i < g_stats.size(); ++i) if (strcmp(g_stats[i].name, name) == 0) return &g_stats[i];
    LqKernelStat s; s.name = name; s.ms = 0; s.launches = 0; s.bytes = 0; g_stats.push_back(s); return &g_stats.back();
}
}

void lq_prof_enable(int on) { g_on = on != 0; }
int lq_prof_on() { return g_on; }
void lq_prof_begin(const char *name, cudaStream_t st) { if (!g_on) return; g_open = name; g_e0 = get_event(); cudaEventRecord(g_e0, st); }
void lq_prof_end(cudaStream_t st, uint64_t launches, uint64_t bytes)
{
    g_launches += launches;
    if (!g_on || !g_open) return;
    Pending p; p.name = g_open; p.e0 = g_e0; p.e1 = get_event(); p.launches = launches; p.bytes = bytes;
    cudaEventRecord(p.e1, st);
    g_pending.push_back(p); g_open = 0;
}
void lq_prof_count_launch(uint64_t n) { g_launches += n; }
void lq_prof_add_bytes(const char *name, uint64_t bytes) { if (g_on && bytes) slot(name)->bytes += bytes; }
void lq_prof_h2d(uint64_t b) { g_h2d += b; }
void lq_prof_d2h(uint64_t b) { g_d2h += b; }
void lq_prof_collect()
{
    for (size_t i = 0; i < g_pending.size(); ++i) {
        Pending &p = g_pending[i];
        float ms = 0.f;
        if (cudaEventSynchronize(p.e1) == cudaSuccess && cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) {
            LqKernelStat *s = slot(p.name); s->ms += ms; s->launches += p.launches; s->bytes += p.bytes;
        }
        g_pool.push_back(p.e0); g_pool.push_back(p.e1);
    }
    g_pending.clear();
}

extern "C" void lqcov_profile_enable(int on) { lq_prof_enable(on); }
extern "C" void lqcov_profile_reset(void) { lq_prof_collect(); g_stats.clear(); g_launches = g_h2d = g_d2h = 0; }
/* JSON: {"launches":L,"h2d_bytes":..,"d2h_bytes":..,"kernels":[{"name":..,"ms":..,"launches":..,"bytes":..},...]} ; returns needed length */
extern "C" size_t lqcov_profile_json(char *buf, size_t cap)
{
    lq_prof_collect();
    std::string s = "{\"launches\":" + std::to_string(g_launches) + ",\"h2d_bytes\":" + std::to_string(g_h2d) + ",\"d2h_bytes\":" + std::to_string(g_d2h) + ",\"kernels\":[";
    for (size_t i = 0; i < g_stats.size(); ++i) {
        char t[256];
        snprintf(t, sizeof t, "%s{\"name\":\"%s\",\"ms\":%.6f,\"launches\":%llu,\"bytes\":%llu}", i ? "," : "", g_stats[i].name, g_stats[i].ms,
                 (unsigned long long)g_stats[i].launches, (unsigned long long)g_stats[i].bytes);
        s += t;
    }
    s += "]}";
    if (buf && cap) { size_t n = s.size() < cap - 1 ? s.size() : cap - 1; memcpy(buf, s.data(), n); buf[n] = 0; }
    return s.size() + 1;
}
