// TEST INFRASTRUCTURE.  The kernel SOURCES of the library -- K3 (rxm_k3.cu), K2 (rxm_k2.cu), K1B
// (rxm_k1b.cu) and the tokeniser (rxm_tok.cu), kernels AND launch functions -- compiled for the
// HOST under the SIMT emulator of simt_shim.hpp, so that the code the GPU runs (not a restatement
// of it) is checked against the golden vectors in the CPU test tier, and a divergent collective,
// a deadlock or an endless loop is reported here instead of hanging a GPU.  (K1's five PTX helpers --
// cp.async, ld.shared.v4, mad.lo -- have host twins in rxm_k1.cu; everything else is the same text.)
// Built into tests/hostsim/libhostsim.so and loaded
// only by tests.
#define RXM_SIMT_HOST 1
#include "simt_shim.hpp"

#include <string>
#include <vector>

static unsigned long long rxm_k3_simt_iterations = 0;  // counted by the K3 kernel under RXM_SIMT_HOST
#include "../../re2-modification_b200/csrc/rxm_k1.cu"
#include "../../re2-modification_b200/csrc/rxm_k1b.cu"
#include "../../re2-modification_b200/csrc/rxm_k2.cu"
#include "../../re2-modification_b200/csrc/rxm_k3.cu"
#include "../../re2-modification_b200/csrc/rxm_k4.cu"
#include "../../re2-modification_b200/csrc/rxm_tok.cu"

namespace {

struct Run {  // one emulated API call: launch defaults in, failure report out
    Run(uint64_t limit, uint64_t seed) {
        simt::State &s = simt::S();
        s.default_limit = limit;
        s.default_seed = seed;
        s.last_failure = 0;
        s.last_msg[0] = 0;
        s.launches = 0;
    }
    int finish(int st, char *msg_out, uint32_t msg_cap) const {
        const simt::State &s = simt::S();
        if (msg_out && msg_cap) {
            strncpy(msg_out, s.last_msg, msg_cap - 1);
            msg_out[msg_cap - 1] = 0;
        }
        if (s.last_failure) return 100 + s.last_failure;
        return st;
    }
};

struct DevTables {  // what rxm_api.cu uploads for the MFA engines
    std::vector<uint16_t> eb;
    std::vector<uint64_t> er;
    rxm::MfaView view(const rxm_tables *t) {
        eb.resize(t->n_states + 1);
        for (uint32_t q = 0; q <= t->n_states; q++) eb[q] = uint16_t(t->edge_begin[q]);
        er.resize(t->n_edges);
        for (uint32_t e = 0; e < t->n_edges; e++)
            er[e] = rxm::pack_edge(t->edge_kind[e], t->edge_sym[e], t->edge_to[e], t->edge_open[e], t->edge_close[e]);
        return rxm::MfaView{eb.data(), er.data(), t->n_states, t->start, t->finish, t->reversed};
    }
};

std::vector<rxm::K1Rec> make_recs(const uint32_t *order_idx, uint64_t n) {
    std::vector<rxm::K1Rec> recs;
    if (order_idx) {
        recs.resize(n);
        for (uint64_t i = 0; i < n; i++) recs[i] = rxm::K1Rec{0, 0, order_idx[i]};
    }
    return recs;
}

}  // namespace

// K3 through rxm::k3_launch (blocks of 8 warps, as on the device) with `tile` lanes per string
// (8, 16, 32; the launch function may widen it).  order_idx: null (strings handed out by index) or
// a permutation standing for the tile sort's order.  Two blocks are launched.  Returns 0; an RXM
// status; or 100 + the emulator's failure code (1 divergent collective, 2 deadlock, 3 budget)
// with the report in msg_out.  seed: 0 = threads run round-robin between collectives, else shuffled.
extern "C" int hostsim_k3_batch(const rxm_tables *t, const uint8_t *chars, const uint64_t *off, uint64_t n,
                                uint8_t *out, uint32_t tile, const uint32_t *order_idx, uint64_t limit,
                                unsigned long long *overflow_out, char *msg_out, uint32_t msg_cap, uint64_t seed) {
    rxm::MfaProgram prog;
    std::string err;
    int st = rxm::compile_programs(*t, prog, &err);
    if (st != RXM_OK) return st;
    DevTables dt;
    const rxm::MfaView v = dt.view(t);
    rxm::ProgView gp{prog.items.data(), prog.begin.data(), prog.count.data(), prog.n_cells};
    std::vector<rxm::K1Rec> recs = make_recs(order_idx, n);
    unsigned long long work[2] = {0, 0};  // [0] overflow, [1] ticket counter (as rxm_api.cu lays them out)
    Run run(limit, seed);
    int launched = 0;
    st = rxm::k3_launch(v, gp, uint32_t(prog.items.size()), uint32_t(prog.begin.size()), t->n_cells, tile, chars,
                        rxm::Spans{off, off + 1}, order_idx ? recs.data() : nullptr, n, out, &work[0], &work[1],
                        /*sm_count=*/2, /*sharing=*/1, nullptr, &launched);
    if (overflow_out) *overflow_out = work[0];
    return run.finish(st, msg_out, msg_cap);
}

// rxm_k4.cu: k4_coop_verify (phase A's verification by the whole warp) on its own: case k asks for the first
// j in [vp[k], vcap[k]) with s[j] != s[j - delta[k]] in the string s = chars[beg[k], beg[k] + len[k]).
namespace {
__global__ void coop_verify_kernel(const uint8_t *chars, const uint64_t *beg, const uint32_t *len, const uint32_t *delta,
                                   const uint32_t *vp, const uint32_t *vcap, uint32_t cases, uint32_t *result) {
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t k = 0; k < cases; k++) {
        const uint32_t r = rxm::k4_coop_verify(chars + beg[k], len[k], delta[k], vp[k], vcap[k], lane);
        if (lane == (k & 31u)) result[k] = r;
    }
}
}  // namespace
extern "C" int hostsim_coop_verify(const uint8_t *chars, const uint64_t *beg, const uint32_t *len, const uint32_t *delta,
                                   const uint32_t *vp, const uint32_t *vcap, uint32_t cases, uint32_t *result,
                                   char *msg_out, uint32_t msg_cap, uint64_t seed) {
    Run run(2000000000ull, seed);
    RXM_LAUNCH(coop_verify_kernel, 1, 32, 0, nullptr, chars, beg, len, delta, vp, vcap, cases, result);
    return run.finish(0, msg_out, msg_cap);
}

// K4 through rxm::k4_launch (128-thread blocks, per-thread sets of `maxl` slots in emulated shared
// memory), then -- as rxm_api.cu does -- K3 over the strings K4 handed on (the redo list) when the
// automaton has more nodes than a thread has slots.  redo_out (may be null) <- strings handed on.
extern "C" int hostsim_k4_batch(const rxm_tables *t, const uint8_t *chars, const uint64_t *off, uint64_t n,
                                uint8_t *out, uint32_t maxl, const uint32_t *order_idx, uint64_t limit,
                                unsigned long long *overflow_out, unsigned long long *redo_out, char *msg_out,
                                uint32_t msg_cap, uint64_t seed) {
    rxm::MfaProgram prog;
    std::string err;
    int st = rxm::compile_programs(*t, prog, &err);
    if (st != RXM_OK) return st;
    DevTables dt;
    const rxm::MfaView v = dt.view(t);
    const std::vector<uint32_t> k4_lists = rxm::k4_pack_lists(prog);
    rxm::K4Prog kp{prog.items.data(), k4_lists.data(), prog.sel.data(), prog.n_cells, prog.n_classes,
                   uint32_t(prog.begin.size())};
    std::vector<rxm::K1Rec> recs = make_recs(order_idx, n);
    // the tile sort's layout: a short last tile is padded out by the kernel's own skip rule, so the
    // records of a permutation must be laid out tile by tile as k1_tilesort_kernel does
    unsigned long long work[2] = {0, 0};
    std::vector<uint32_t> redo(n + 1);
    unsigned long long redo_n = 0;
    if (maxl == 0 || maxl > rxm::K4_POOL_MAX) maxl = rxm::k4_pool_for(t->n_states);  // the planner's choice
    const bool use_redo = 2 * t->n_states > maxl;  // the current set and the one being built share the pool
    Run run(limit, seed);
    int launched = 0;
    st = rxm::k4_launch(v, kp, uint32_t(prog.items.size()), uint32_t(prog.begin.size()), uint32_t(prog.sel.size()),
                        t->n_cells, maxl, chars, rxm::Spans{off, off + 1}, order_idx ? recs.data() : nullptr, n, out,
                        &work[0], &work[1], use_redo ? redo.data() : nullptr, &redo_n, /*sm_count=*/2, /*sharing=*/1,
                        nullptr, &launched);
    if (st == RXM_OK && simt::S().last_failure == 0 && use_redo) {
        rxm::ProgView gp{prog.items.data(), prog.begin.data(), prog.count.data(), prog.n_cells};
        const uint32_t tile = prog.max_count <= 8 ? 8 : (prog.max_count <= 16 ? 16 : 32);
        st = rxm::k3_launch(v, gp, uint32_t(prog.items.size()), uint32_t(prog.begin.size()), t->n_cells, tile, chars,
                            rxm::Spans{off, off + 1}, nullptr, n, out, &work[0], &work[1], 2, 1, nullptr, &launched,
                            redo.data(), &redo_n);
    }
    if (overflow_out) *overflow_out = work[0];
    if (redo_out) *redo_out = redo_n;
    return run.finish(st, msg_out, msg_cap);
}

// K1 through rxm::plan_dfa, rxm::k1_build_tables and rxm::k1_launch: the tile sort, then the scan
// kernel the tables select (quad stride / direct / two-lookup; hostsim_k1_no_quad stands for RXM_OPT_K1_NO_QUAD, RXM_K1_VARIANT is
// read as on the device).  info3 (may be null) <- sets, byte classes, bytes per lookup.
static bool g_k1_no_quad = false, g_k1_no_oct = false;  // RXM_OPT_K1_NO_QUAD / _NO_OCT for the next hostsim_k1_batch calls
extern "C" void hostsim_k1_no_quad(int on) { g_k1_no_quad = (on & 1) != 0; g_k1_no_oct = (on & 2) != 0; }
extern "C" int hostsim_k1_batch(const rxm_tables *t, const uint8_t *chars, const uint64_t *off, uint64_t n,
                                uint8_t *out, uint64_t limit, unsigned long long *overflow_out, uint32_t *info3,
                                char *msg_out, uint32_t msg_cap, uint64_t seed) {
    rxm::DfaPlan p;
    std::string err;
    int st = rxm::plan_dfa(*t, p, &err);
    if (st != RXM_OK) return st;
    rxm::K1Tables kt;
    std::vector<uint8_t> table, accept;
    st = rxm::k1_build_tables(p, kt, table, accept, &err, g_k1_no_quad, g_k1_no_oct);
    if (st != RXM_OK) return st;
    if (info3) {
        info3[0] = p.n_states;
        info3[1] = p.n_classes;
        info3[2] = kt.quad == 2 ? 8 : (kt.quad ? 4 : 1);
    }
    table.resize(table.size() + 64);  // the kernels copy whole 16-byte vectors
    accept.resize(accept.size() + 256);
    std::vector<rxm::K1Rec> recs(n + (n >> 3) + 32);
    uint32_t counter[64] = {0};
    unsigned long long overflow = 0;
    Run run(limit, seed);
    int launched = 0;
    rxm::K1Launch a{table.data(), accept.data(), chars, rxm::Spans{off, off + 1}, n, out, recs.data(), counter, &overflow,
                    /*sm_count=*/2, nullptr};
    st = rxm::k1_launch(kt, a, &launched);
    if (overflow_out) *overflow_out = overflow;
    return run.finish(st, msg_out, msg_cap);
}

// K2 through rxm::k2_launch (128-thread blocks, state in emulated shared or local memory).
extern "C" int hostsim_k2_batch(const rxm_tables *t, const uint8_t *chars, const uint64_t *off, uint64_t n,
                                uint8_t *out, uint64_t limit, unsigned long long *overflow_out, char *msg_out,
                                uint32_t msg_cap, uint64_t seed) {
    DevTables dt;
    const rxm::MfaView v = dt.view(t);
    unsigned long long work[2] = {0, 0};
    Run run(limit, seed);
    int launched = 0;
    const int st = rxm::k2_launch(v, t->n_cells, t->n_edges, chars, rxm::Spans{off, off + 1}, n, out, &work[0], &work[1], 2,
                                  nullptr, &launched);
    if (overflow_out) *overflow_out = work[0];
    return run.finish(st, msg_out, msg_cap);
}

// K1B through rxm::k1b_mask_launch (mode 0: follow masks; returns 1 if the table does not satisfy
// their structural condition) or rxm::k1b_launch (mode 1: edge walk).
extern "C" int hostsim_k1b_batch(const rxm_tables *t, const uint8_t *chars, const uint64_t *off, uint64_t n,
                                 uint8_t *out, int mode, const uint32_t *order_idx, uint64_t limit,
                                 unsigned long long *overflow_out, char *msg_out, uint32_t msg_cap, uint64_t seed) {
    std::string err;
    int st = rxm::check_nfa_bitset(*t, &err);
    if (st != RXM_OK) return st;
    std::vector<rxm::K1Rec> recs = make_recs(order_idx, n);
    const rxm::K1Rec *rp = order_idx ? recs.data() : nullptr;
    unsigned long long work[2] = {0, 0};
    Run run(limit, seed);
    int launched = 0;
    if (mode == 0) {
        rxm::BitsetMasks bm;
        rxm::plan_bitset_masks(*t, bm);
        if (!bm.ok) return 1;
        st = rxm::k1b_mask_launch(bm.ls.data(), bm.byte_class, t->n_states, bm.n_classes, t->start, bm.accept[0],
                                  bm.accept[1], t->reversed, chars, rxm::Spans{off, off + 1}, rp, n, out, &work[0], &work[1],
                                  2, nullptr, &launched);
    } else {
        std::vector<uint16_t> eb;
        std::vector<uint32_t> ed;
        rxm::k1b_build_tables(*t, eb, ed);
        st = rxm::k1b_launch(eb.data(), ed.data(), t->n_states, t->n_edges, t->start, t->finish, t->reversed, chars,
                             rxm::Spans{off, off + 1}, rp, n, out, &work[0], &work[1], 2, nullptr, &launched);
    }
    if (overflow_out) *overflow_out = work[0];
    return run.finish(st, msg_out, msg_cap);
}

// The tokeniser through rxm::tok_launch: text[0, nbytes) (the buffer must be readable 16 bytes to
// either side) -> begin / end of the first `cap` tokens, result[0] = tokens found, result[1] =
// index of the first token "exit" (~0 if none).
extern "C" int hostsim_tok(const uint8_t *text, uint64_t nbytes, uint64_t *begin, uint64_t *end, uint64_t cap,
                           unsigned long long *result, uint64_t limit, char *msg_out, uint32_t msg_cap, uint64_t seed) {
    const uint64_t blocks = rxm::tok_blocks(nbytes);
    std::vector<uint64_t> masks(blocks * 256 + 8, 0xcdcdcdcdcdcdcdcdull), counts(blocks * 9 + 8, 0xcdcdcdcdcdcdcdcdull);
    rxm::TokWork w{masks.data(), counts.data(), blocks, result};
    Run run(limit, seed);
    int launched = 0;
    const int st = rxm::tok_launch(text, nbytes, begin, end, cap, w, /*sm_count=*/2, nullptr, &launched);
    return run.finish(st, msg_out, msg_cap);
}

// Collectives the last emulated block executed; lock-step iterations of K3's warps since the last call.
extern "C" unsigned long long hostsim_simt_collectives() { return simt::S().collectives; }
extern "C" unsigned long long hostsim_k3_iterations() {
    const unsigned long long r = rxm_k3_simt_iterations;
    rxm_k3_simt_iterations = 0;
    return r;
}

// The emulator's own checks on tiny kernels: 0 well-formed (returns 0 when the results are right),
// 1 vote in a lane-dependent branch, 2 a lane returns early, 3 short-circuited votes, 4 endless
// loop, 5 partial mask; 6 / 7 an unsynchronised neighbour read under round-robin / shuffled lane
// order (returns 1000 + the number of lanes that saw the neighbour's write); 8 a four-warp block
// with __syncthreads and static shared memory (0 when right); 9 __syncthreads at two call sites.
// Otherwise returns the emulator's failure code.
extern "C" int hostsim_simt_selftest(int which, char *msg_out, uint32_t msg_cap) {
    const char *msg = "";
    unsigned long long race_hits = 0;  // lanes that saw their neighbour's write
    bool ok = true;
    int rc;
    if (which >= 8) {
        Run run(100000, 0);
        unsigned total[2] = {0, 0};
        auto kernel = [&]() {
            __shared__ unsigned s_part[4];
            const unsigned t = threadIdx.x, lane = t & 31u, wp = t >> 5;
            const unsigned sum = __reduce_add_sync(0xffffffffu, t + blockIdx.x);
            if (lane == 0) s_part[wp] = sum;
            __syncthreads();
            if (t == 0) total[blockIdx.x] = s_part[0] + s_part[1] + s_part[2] + s_part[3];
            if (which == 9) {
                if (t & 64) __syncthreads();
                else __syncthreads();
            }
        };
        simt::launch_grid(2, 128, 0, kernel);
        rc = simt::S().last_failure;
        msg = simt::S().last_msg;
        if (which == 8 && rc == 0 && (total[0] != 127 * 64 || total[1] != 127 * 64 + 128)) rc = -1;
    } else {
        rc = simt::launch_warp(
            [&]() {
                const uint32_t lane = threadIdx.x;
                if (which == 0) {
                    uint32_t x = lane + 1;
                    for (int d = 16; d >= 1; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
                    const uint32_t up = __shfl_up_sync(0xffffffffu, lane, 1, 8);
                    const uint32_t pick = __shfl_sync(0xffffffffu, lane, 3, 8);
                    const uint32_t b = __ballot_sync(0xffffffffu, lane & 1);
                    const uint32_t mn = __reduce_min_sync(0xffffffffu, lane + 5), mx = __reduce_max_sync(0xffffffffu, lane);
                    __syncwarp(0xffffffffu);
                    if (x != 528 || up != ((lane & 7) ? lane - 1 : lane) || pick != (lane & ~7u) + 3 || b != 0xaaaaaaaau ||
                        mn != 5 || mx != 31)
                        ok = false;
                } else if (which == 1) {
                    if (lane < 16) (void)__ballot_sync(0xffffffffu, true);
                    else (void)__any_sync(0xffffffffu, true);
                } else if (which == 2) {
                    if (lane == 7) return;
                    (void)__ballot_sync(0xffffffffu, true);
                } else if (which == 3) {
                    const bool r = __all_sync(0xffffffffu, true) && (lane < 8 || __any_sync(0xffffffffu, true));
                    (void)r;
                    (void)__ballot_sync(0xffffffffu, true);
                } else if (which == 4) {
                    for (;;) (void)__ballot_sync(0xffffffffu, true);
                } else if (which == 5) {
                    (void)__ballot_sync(0x0000ffffu, true);
                } else {  // 6 / 7: a read of the neighbour's word with no __syncwarp between write and read
                    uint32_t *w = reinterpret_cast<uint32_t *>(simt::S().smem);
                    (void)__ballot_sync(0xffffffffu, true);
                    w[lane] = lane + 100;
                    if (w[(lane + 1) & 31] == ((lane + 1) & 31) + 100) atomicAdd(&race_hits, 1ull);
                    (void)__ballot_sync(0xffffffffu, true);
                }
            },
            256, 100000, &msg, which == 7 ? 12345 : 0);
    }
    if (msg_out && msg_cap) {
        strncpy(msg_out, msg, msg_cap - 1);
        msg_out[msg_cap - 1] = 0;
    }
    if (which == 0 && rc == 0 && !ok) return -1;
    if ((which == 6 || which == 7) && rc == 0) return 1000 + int(race_hits);
    return rc;
}
