// TEST INFRASTRUCTURE.  Compiles the kernels' shared simulation core
// (re2-modification_b200/csrc/rxm_mfa_core.cuh) for the HOST so that its logic
// can be compared with the oracle in the CPU test tier, before GPU time is
// spent.  It is built into tests/hostsim/libhostsim.so and loaded only by
// tests; librxm.so does not contain or call it.
#include <cstdint>
#include <vector>

#include "../../include/rxm.h"

// How periodic is the work?  Before every step ProgSim runs, its starting set is compared with the
// set the previous step started from: `periodic` = the same set with every `first` moved on by the
// distance between the two steps (cells, flags, stamps identical), the same input letter, forward
// reading -- a step whose outcome is fixed by the outcome of its block compares alone;
// `confirmed` = the step before it was periodic too, over the same distance (the phase goes on).
static uint64_t g_obs_steps = 0, g_obs_periodic = 0, g_obs_confirmed = 0;
static int g_obs_relaxed = 0;  // 1: open cells may grow by the distance (letter loops inside a cell)
static int g_obs_period = 1;   // compare with the set this many steps back (1..8)
template <class CfgT, class ReaderT>
static void progsim_observe(const CfgT *cur, uint32_t m, uint32_t i, const ReaderT &rd);
#define RXM_PROGSIM_OBSERVE(cur, m, i, rd) progsim_observe(cur, m, i, rd)
#include "../../re2-modification_b200/csrc/rxm_mfa_core.cuh"
#include "../../re2-modification_b200/csrc/rxm_k4_core.cuh"
#include "../../re2-modification_b200/csrc/rxm_mfa_dispatch.hpp"
#include "../../re2-modification_b200/csrc/rxm_nfa_core.cuh"
#include "../../re2-modification_b200/csrc/rxm_plan.hpp"

template <class CfgT>
static bool moved_equal(const std::vector<CfgT> &prev, const CfgT *cur, uint32_t m, uint32_t delta) {
    if (m != prev.size()) return false;
    for (uint32_t j = 0; j < m; j++) {
        bool found = false;
        for (const CfgT &q : prev) {
            if (q.node != cur[j].node) continue;
            CfgT moved = q;
            moved.first += delta;
            if (g_obs_relaxed) {  // an open cell may have grown by the letters read in between
                for (int k = 0; k < int(sizeof(q.len) / sizeof(q.len[0])); k++)
                    if (rxm::fl_exists(q.flags, k) && rxm::fl_open(q.flags, k)) {
                        if (moved.len[k] == 0) moved.start[k] = cur[j].start[k];
                        moved.len[k] += delta;
                    }
            }
            found = rxm::cfg_same(moved, cur[j]) && q.born == cur[j].born;
            break;
        }
        if (!found) return false;
    }
    return true;
}

// g_obs_period = p: the set is compared with the one p steps back (p = 1: the step before).
template <class CfgT, class ReaderT>
static void progsim_observe(const CfgT *cur, uint32_t m, uint32_t i, const ReaderT &rd) {
    constexpr int H = 8;
    static std::vector<CfgT> hist[H];
    static uint32_t hist_i[H];
    static int n_hist = 0;
    static uint32_t prev_delta = 0;
    static bool prev_periodic = false;
    if (i == 0) n_hist = 0, prev_periodic = false;
    g_obs_steps++;
    bool periodic = false;
    uint32_t delta = 0;
    const int p = g_obs_period;
    if (n_hist >= p) {
        const int slot = (n_hist - p) % H;
        delta = i - hist_i[slot];
        periodic = !rd.reversed && i < rd.n && rd.at(i) == rd.at(hist_i[slot]) && moved_equal(hist[slot], cur, m, delta);
        if (periodic) {
            g_obs_periodic++;
            if (prev_periodic && prev_delta == delta) g_obs_confirmed++;
        }
    }
    hist[n_hist % H].assign(cur, cur + m);
    hist_i[n_hist % H] = i;
    n_hist++;
    prev_delta = delta;
    prev_periodic = periodic;
}
extern "C" void hostsim_periodic_relaxed(int on) { g_obs_relaxed = on; }
extern "C" void hostsim_periodic_period(int p) { g_obs_period = p < 1 ? 1 : (p > 8 ? 8 : p); }
extern "C" void hostsim_periodic_stats(uint64_t *steps, uint64_t *periodic, uint64_t *confirmed) {
    *steps = g_obs_steps;
    *periodic = g_obs_periodic;
    *confirmed = g_obs_confirmed;
    g_obs_steps = g_obs_periodic = g_obs_confirmed = 0;
}

static uint64_t g_steps_run = 0, g_steps_skipped = 0, g_steps_replayed = 0;
static int g_prog_replay = 0;  // ProgSim::replay for the next hostsim_prog_batch calls
extern "C" void hostsim_step_stats(uint64_t *run, uint64_t *skipped) {
    *run = g_steps_run;
    *skipped = g_steps_skipped;
    g_steps_run = g_steps_skipped = 0;
}
// Repeated steps (ProgSim::replay): switch, and the number of steps answered by their block compares
// alone since the last call.
extern "C" void hostsim_prog_replay(int on) { g_prog_replay = on; }
extern "C" uint64_t hostsim_steps_replayed() {
    const uint64_t r = g_steps_replayed;
    g_steps_replayed = 0;
    return r;
}

template <int NC, int CAP, int DMAX>
static void run_batch(const rxm::MfaView &v, const uint8_t *chars, const uint64_t *off, uint64_t n,
                      uint8_t *out) {
    auto *sim = new rxm::MfaSim<NC, CAP, DMAX>();
    for (uint64_t i = 0; i < n; i++) {
        rxm::Reader rd{chars + off[i], uint32_t(off[i + 1] - off[i]), v.reversed};
        out[i] = uint8_t(sim->run(v, rd));
        g_steps_run += sim->steps_run;
        g_steps_skipped += sim->steps_skipped;
    }
    delete sim;
}

extern "C" int hostsim_mfa_batch(const rxm_tables *t, const uint8_t *chars, const uint64_t *off,
                                 uint64_t n, uint8_t *out) {
    std::vector<uint16_t> eb(t->n_states + 1);
    for (uint32_t q = 0; q <= t->n_states; q++) eb[q] = uint16_t(t->edge_begin[q]);
    std::vector<uint64_t> er(t->n_edges);
    for (uint32_t e = 0; e < t->n_edges; e++)
        er[e] = rxm::pack_edge(t->edge_kind[e], t->edge_sym[e], t->edge_to[e], t->edge_open[e],
                               t->edge_close[e]);
    rxm::MfaView v{eb.data(), er.data(), t->n_states, t->start, t->finish, t->reversed};
    bool rxm_dispatch_ok = true;
#define CALL(NC, CAP, DMAX) run_batch<NC, CAP, DMAX>(v, chars, off, n, out)
    RXM_MFA_DISPATCH(t->n_cells, t->n_states, CALL);
#undef CALL
    return rxm_dispatch_ok ? 0 : 2;
}

template <int NC, int CAP, int DMAX>
static void run_prog_batch(const rxm::MfaView &v, const rxm::ProgView &pv, const uint8_t *chars,
                           const uint64_t *off, uint64_t n, uint8_t *out) {
    auto *sim = new rxm::ProgSim<NC, CAP>();
    sim->replay = g_prog_replay != 0;
    for (uint64_t i = 0; i < n; i++) {
        rxm::Reader rd{chars + off[i], uint32_t(off[i + 1] - off[i]), v.reversed};
        out[i] = uint8_t(sim->run(v, pv, rd));
        g_steps_run += sim->steps_run;
        g_steps_skipped += sim->steps_skipped;
        g_steps_replayed += sim->steps_replayed;
    }
    delete sim;
}

// Edge programs (rxm_plan.cpp: compile_programs) run by the sequential interpreter.
extern "C" int hostsim_prog_batch(const rxm_tables *t, const uint8_t *chars, const uint64_t *off,
                                  uint64_t n, uint8_t *out, uint32_t *info2) {
    rxm::MfaProgram prog;
    std::string err;
    int st = rxm::compile_programs(*t, prog, &err);
    if (st != RXM_OK) return st;
    if (info2) {
        info2[0] = uint32_t(prog.items.size());
        info2[1] = prog.max_count;
    }
    std::vector<uint16_t> eb(t->n_states + 1);
    for (uint32_t q = 0; q <= t->n_states; q++) eb[q] = uint16_t(t->edge_begin[q]);
    std::vector<uint64_t> er(t->n_edges);
    for (uint32_t e = 0; e < t->n_edges; e++)
        er[e] = rxm::pack_edge(t->edge_kind[e], t->edge_sym[e], t->edge_to[e], t->edge_open[e],
                               t->edge_close[e]);
    rxm::MfaView v{eb.data(), er.data(), t->n_states, t->start, t->finish, t->reversed};
    rxm::ProgView pv{prog.items.data(), prog.begin.data(), prog.count.data(), prog.n_cells};
    bool rxm_dispatch_ok = true;
#define CALL(NC, CAP, DMAX) run_prog_batch<NC, CAP, DMAX>(v, pv, chars, off, n, out)
    RXM_MFA_DISPATCH(t->n_cells, t->n_states, CALL);
#undef CALL
    return rxm_dispatch_ok ? 0 : 2;
}

extern "C" int hostsim_span_equal(const uint8_t *a, const uint8_t *b, uint32_t L) { return rxm::k4_span_equal(a, b, L) ? 1 : 0; }

// K4's per-string simulation (rxm_k4_core.cuh) on the host, plain storage; `maxl` slots per set.
// info3 (may be null) <- steps run in full, repeated steps answered by their compares, strings that met a limit.
template <int NC>
static void run_k4_batch(const rxm::MfaView &v, const rxm::K4Prog &kp, uint32_t maxl, const uint8_t *chars,
                         const uint64_t *off, uint64_t n, uint8_t *out, uint64_t *info3) {
    std::vector<uint32_t> words(rxm::k4_words(NC, maxl));
    rxm::K4Sim<NC, 1> sim;
    sim.base = words.data();
    sim.pool = maxl;
    sim.use_map = v.n_states <= rxm::K4_MAP_STATES && maxl <= 15u;
    for (uint64_t i = 0; i < n; i++) {
        const int r = sim.run(v, kp, chars + off[i], uint32_t(off[i + 1] - off[i]));
        out[i] = uint8_t(r);
        if (info3) {
            info3[0] += sim.steps_run;
            info3[1] += sim.steps_replayed;
            info3[2] += (r == 2);
        }
    }
}

extern "C" int hostsim_k4core_batch(const rxm_tables *t, const uint8_t *chars, const uint64_t *off, uint64_t n,
                                    uint8_t *out, uint32_t maxl, uint64_t *info3) {
    rxm::MfaProgram prog;
    std::string err;
    int st = rxm::compile_programs(*t, prog, &err);
    if (st != RXM_OK) return st;
    std::vector<uint16_t> eb(t->n_states + 1);
    for (uint32_t q = 0; q <= t->n_states; q++) eb[q] = uint16_t(t->edge_begin[q]);
    std::vector<uint64_t> er(t->n_edges);
    for (uint32_t e = 0; e < t->n_edges; e++)
        er[e] = rxm::pack_edge(t->edge_kind[e], t->edge_sym[e], t->edge_to[e], t->edge_open[e], t->edge_close[e]);
    rxm::MfaView v{eb.data(), er.data(), t->n_states, t->start, t->finish, t->reversed};
    const std::vector<uint32_t> k4_lists = rxm::k4_pack_lists(prog);
    rxm::K4Prog kp{prog.items.data(), k4_lists.data(), prog.sel.data(), prog.n_cells, prog.n_classes,
                   uint32_t(prog.begin.size())};
    if (info3) info3[0] = info3[1] = info3[2] = 0;
    if (maxl == 0 || maxl > rxm::K4_POOL_MAX) maxl = rxm::k4_pool_for(t->n_states);  // the planner's choice
    if (t->n_cells <= 1) run_k4_batch<1>(v, kp, maxl, chars, off, n, out, info3);
    else if (t->n_cells <= 2) run_k4_batch<2>(v, kp, maxl, chars, off, n, out, info3);
    else run_k4_batch<4>(v, kp, maxl, chars, off, n, out, info3);
    return 0;
}

// DFA built by the planner, stepped on the host (checks plan_dfa against the oracle).
extern "C" int hostsim_dfa_batch(const rxm_tables *t, const uint8_t *chars, const uint64_t *off,
                                 uint64_t n, uint8_t *out, uint32_t *info3) {
    rxm::DfaPlan p;
    std::string err;
    int st = rxm::plan_dfa(*t, p, &err);
    if (st != RXM_OK) return st;
    if (info3) {
        info3[0] = p.n_states;
        info3[1] = p.n_classes;
        info3[2] = p.exact_step_differs;
    }
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t *s = chars + off[i];
        const uint64_t len = off[i + 1] - off[i];
        uint32_t q = p.start;
        for (uint64_t j = 0; j < len && q; j++) {
            const uint8_t b = p.reversed ? s[len - 1 - j] : s[j];
            q = p.trans[size_t(p.byte_class[b]) * p.n_states + q];
        }
        out[i] = p.accept[q];
    }
    return 0;
}

// K1B's bit-parallel step from follow masks on the host.  Returns 0, 1 if the structural
// condition for the masks does not hold for this table (the caller then expects the walk), or the
// planner's status.
extern "C" int hostsim_nfa_mask_batch(const rxm_tables *t, const uint8_t *chars, const uint64_t *off,
                                      uint64_t n, uint8_t *out) {
    std::string err;
    int st = rxm::check_nfa_bitset(*t, &err);
    if (st != RXM_OK) return st;
    rxm::BitsetMasks bm;
    rxm::plan_bitset_masks(*t, bm);
    if (!bm.ok) return 1;
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t *s = chars + off[i];
        const uint32_t len = uint32_t(off[i + 1] - off[i]);
        rxm::Bits128 S{0, 0};
        S.set(t->start);
        for (uint32_t k = 0; k < len && !S.empty(); k++) {
            const uint8_t b = t->reversed ? s[len - 1 - k] : s[k];
            S = rxm::nfa_mask_step(bm.ls.data() + size_t(bm.byte_class[b]) * t->n_states * 2, S);
        }
        out[i] = ((S.lo & bm.accept[0]) | (S.hi & bm.accept[1])) ? 1 : 0;
    }
    return 0;
}

// The K1B step (rxm_nfa_core.cuh) on the host: returns 0, the planner's status if the bit-set
// engine's static checks reject the table, or 2 if a string overflowed the recursion stack.
extern "C" int hostsim_nfa_bits_batch(const rxm_tables *t, const uint8_t *chars, const uint64_t *off,
                                      uint64_t n, uint8_t *out) {
    std::string err;
    int st = rxm::check_nfa_bitset(*t, &err);
    if (st != RXM_OK) return st;
    std::vector<uint16_t> eb(t->n_states + 1);
    std::vector<uint32_t> ed(t->n_edges);
    for (uint32_t q = 0; q <= t->n_states; q++) eb[q] = uint16_t(t->edge_begin[q]);
    for (uint32_t e = 0; e < t->n_edges; e++) ed[e] = rxm::nfa_pack_edge(t->edge_kind[e], t->edge_sym[e], t->edge_to[e]);
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t *s = chars + off[i];
        const uint32_t len = uint32_t(off[i + 1] - off[i]);
        rxm::Bits128 S{0, 0}, N{0, 0};
        S.set(t->start);
        bool ok = true;
        for (uint32_t k = 0; k < len && ok; k++) {
            ok = rxm::nfa_bits_step(eb.data(), ed.data(), t->finish, S, t->reversed ? s[len - 1 - k] : s[k], N);
            S = N;
            if (S.empty()) break;
        }
        if (ok) ok = rxm::nfa_bits_step(eb.data(), ed.data(), t->finish, S, -1, N);
        if (!ok) return 2;
        out[i] = N.test(t->finish) ? 1 : 0;
    }
    return 0;
}
