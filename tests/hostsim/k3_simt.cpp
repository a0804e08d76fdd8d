// TEST INFRASTRUCTURE.  The K3 kernel source (re2-modification_b200/csrc/rxm_k3.cu) compiled
// for the HOST under the single-warp SIMT emulator of simt_shim.hpp, so that the kernel itself
// -- not only the simulation core it shares with K2 -- is checked against the golden vectors in
// the CPU test tier, and a divergent collective or an endless loop is reported here instead of
// hanging a GPU.  Built into tests/hostsim/libhostsim.so and loaded only by tests.
#define RXM_SIMT_HOST 1
#include "simt_shim.hpp"

#include <string>
#include <vector>

static unsigned long long rxm_k3_simt_iterations = 0;  // counted by the kernel under RXM_SIMT_HOST
#include "../../re2-modification_b200/csrc/rxm_k3.cu"

namespace {

template <int NC, int TILE>
int run_k3(const rxm::MfaView &v, const rxm::ProgView &gp, uint32_t n_items, uint32_t n_keys, const uint8_t *chars,
           const uint64_t *off, const rxm::K1Rec *recs, uint64_t n, uint8_t *out, unsigned long long *overflow,
           uint64_t limit, const char **msg, uint64_t seed) {
    // shared-memory budget exactly as launch_k3 computes it, for one warp
    constexpr int TILES = 32 / TILE;
    const uint32_t SP = (v.n_states + TILE - 1u) & ~uint32_t(TILE - 1);
    const size_t per_tile = size_t(SP) * 2 * (8 + 4 + 8 * NC) + size_t(SP) * 12 + 16;
    const size_t tab = (size_t(n_keys) * 8 + 15) & ~size_t(15);
    const size_t smem = tab + size_t(n_items) * sizeof(rxm::ProgItem) + TILES * per_tile;
    unsigned long long next = 0;
    rxm::Spans sp{off, off + 1};
    return simt::launch_warp(
        [&]() {
            rxm::k3_mfa_warp_kernel<NC, TILE>(v, gp, n_items, n_keys, 1u, chars, sp, recs, n, out, overflow, &next);
        },
        smem, limit, msg, seed);
}

template <int NC>
int run_k3_tile(uint32_t tile, const rxm::MfaView &v, const rxm::ProgView &gp, uint32_t n_items, uint32_t n_keys,
                const uint8_t *chars, const uint64_t *off, const rxm::K1Rec *recs, uint64_t n, uint8_t *out,
                unsigned long long *overflow, uint64_t limit, const char **msg, uint64_t seed) {
    if (tile == 8) return run_k3<NC, 8>(v, gp, n_items, n_keys, chars, off, recs, n, out, overflow, limit, msg, seed);
    if (tile == 16) return run_k3<NC, 16>(v, gp, n_items, n_keys, chars, off, recs, n, out, overflow, limit, msg, seed);
    return run_k3<NC, 32>(v, gp, n_items, n_keys, chars, off, recs, n, out, overflow, limit, msg, seed);
}

}  // namespace

// Runs the batch through the emulated K3 kernel with `tile` lanes per string (8, 16, 32).
// order: null (strings handed out by index) or the K1Rec array of the tile sort's order (only
// .idx is read).  Returns 0; RXM status if the programs do not compile; 100 + emulator failure
// code (1 divergent collective, 2 deadlock, 3 watchdog) with the report in msg_out.
// seed: 0 = lanes run round-robin between collectives, else a shuffled order (races between lanes).
extern "C" int hostsim_k3_batch(const rxm_tables *t, const uint8_t *chars, const uint64_t *off, uint64_t n,
                                uint8_t *out, uint32_t tile, const uint32_t *order_idx, uint64_t limit,
                                unsigned long long *overflow_out, char *msg_out, uint32_t msg_cap, uint64_t seed) {
    rxm::MfaProgram prog;
    std::string err;
    int st = rxm::compile_programs(*t, prog, &err);
    if (st != RXM_OK) return st;
    if (prog.n_cells > 4) return RXM_ERR_UNSUPPORTED;
    std::vector<uint16_t> eb(t->n_states + 1);
    for (uint32_t q = 0; q <= t->n_states; q++) eb[q] = uint16_t(t->edge_begin[q]);
    std::vector<uint64_t> er(t->n_edges);
    for (uint32_t e = 0; e < t->n_edges; e++)
        er[e] = rxm::pack_edge(t->edge_kind[e], t->edge_sym[e], t->edge_to[e], t->edge_open[e], t->edge_close[e]);
    rxm::MfaView v{eb.data(), er.data(), t->n_states, t->start, t->finish, t->reversed};
    rxm::ProgView gp{prog.items.data(), prog.begin.data(), prog.count.data(), prog.n_cells};
    std::vector<rxm::K1Rec> recs;
    if (order_idx) {
        recs.resize(n);
        for (uint64_t i = 0; i < n; i++) recs[i] = rxm::K1Rec{0, 0, order_idx[i]};
    }
    unsigned long long overflow = 0;
    const char *msg = "";
    const uint32_t n_items = uint32_t(prog.items.size()), n_keys = uint32_t(prog.begin.size());
    const rxm::K1Rec *rp = order_idx ? recs.data() : nullptr;
    int rc;
    if (prog.n_cells <= 1) rc = run_k3_tile<1>(tile, v, gp, n_items, n_keys, chars, off, rp, n, out, &overflow, limit, &msg, seed);
    else if (prog.n_cells <= 2) rc = run_k3_tile<2>(tile, v, gp, n_items, n_keys, chars, off, rp, n, out, &overflow, limit, &msg, seed);
    else rc = run_k3_tile<4>(tile, v, gp, n_items, n_keys, chars, off, rp, n, out, &overflow, limit, &msg, seed);
    if (overflow_out) *overflow_out = overflow;
    if (msg_out && msg_cap) {
        strncpy(msg_out, msg, msg_cap - 1);
        msg_out[msg_cap - 1] = 0;
    }
    return rc ? 100 + rc : 0;
}

// The emulator's own checks on tiny kernels: 0 well-formed (returns 0 when the results are right),
// 1 vote in a lane-dependent branch, 2 a lane returns early, 3 short-circuited votes, 4 endless
// loop, 5 partial mask; 6 / 7 an unsynchronised neighbour read under round-robin / shuffled lane
// order (returns 1000 + the number of lanes that saw the neighbour's write).  Otherwise returns the
// emulator's failure code.
extern "C" int hostsim_simt_selftest(int which, char *msg_out, uint32_t msg_cap) {
    const char *msg = "";
    uint32_t sums[32] = {0};
    unsigned long long race_hits = 0;  // lanes that saw their neighbour's write
    bool ok = true;
    const int rc = simt::launch_warp(
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
                sums[lane] = x;
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
                if (lane == 0) sums[0] = 0;
                (void)__ballot_sync(0xffffffffu, true);
                w[lane] = lane + 100;
                if (w[(lane + 1) & 31] == ((lane + 1) & 31) + 100) atomicAdd(&race_hits, 1ull);
                (void)__ballot_sync(0xffffffffu, true);
            }
        },
        256, 100000, &msg, which == 7 ? 12345 : 0);
    if (msg_out && msg_cap) {
        strncpy(msg_out, msg, msg_cap - 1);
        msg_out[msg_cap - 1] = 0;
    }
    if (which == 0 && rc == 0 && !ok) return -1;
    if (which >= 6 && rc == 0) return 1000 + int(race_hits);
    return rc;
}

// Collectives the last emulated launch executed (a proxy for the steps the kernel ran).
extern "C" unsigned long long hostsim_simt_collectives() { return simt::S().collectives; }
// Lock-step iterations (steps of the slowest tile) of all launches since the last call.
extern "C" unsigned long long hostsim_k3_iterations() {
    const unsigned long long r = rxm_k3_simt_iterations;
    rxm_k3_simt_iterations = 0;
    return r;
}
