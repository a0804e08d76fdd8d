// Tuning helper (CPU; not product, not a test): tests/hostsim/hostsim.cpp built with K4's host statistics switched on
// (K4_STAT in rxm_k4_core.cuh) -- configurations, items, fired items, insertions, block compares per step.
// Built and driven by tools/k4_stats.py.
#include <cstdint>
static uint64_t g_find, g_ins, g_reins, g_cmp, g_cmpbytes, g_src, g_active, g_waiting, g_items, g_fire;
#define K4_STAT(x) x
#include "../tests/hostsim/hostsim.cpp"
extern "C" void k4stat_get(uint64_t *o) {
    const uint64_t v[10] = {g_find, g_ins, g_reins, g_cmp, g_cmpbytes, g_src, g_active, g_waiting, g_items, g_fire};
    for (int i = 0; i < 10; i++) o[i] = v[i];
    g_find = g_ins = g_reins = g_cmp = g_cmpbytes = g_src = g_active = g_waiting = g_items = g_fire = 0;
}
