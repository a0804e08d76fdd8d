// K4 -- MFA, ONE THREAD per string, sm_100a.
// Replaces MFA::match (mfa.cpp:215-236) for whole batches.  The simulation is rxm_k4_core.cuh (edge
// programs, repeated steps answered by their block compares, idle steps skipped); this file is the
// kernel around it: tables and per-thread sets in shared memory, strings handed out in the tile
// sort's order so that the 32 strings of a warp have nearly the same length.
//
// Strings that need more than `maxl` configuration slots at a time (the current set and the one being
// built share a thread's pool) are not answered here: their indices go to `redo_list` and K3 runs them.
#include "rxm_k4_core.cuh"
#include "rxm_kernels.cuh"

namespace rxm {

namespace {

constexpr int K4_THREADS = 128;
constexpr uint32_t K4_CHUNK = 64;  // tickets a warp takes from the global counter at a time

template <int NC>
__global__ void __launch_bounds__(K4_THREADS)
k4_mfa_thread_kernel(MfaView v, K4Prog gp, uint32_t n_items, uint32_t n_keys, uint32_t n_sel, uint32_t items_in_smem,
                     uint32_t maxl, const uint8_t *__restrict__ chars, const Spans sp, const K1Rec *__restrict__ recs,
                     uint64_t n, uint8_t *__restrict__ out, unsigned long long *__restrict__ overflow,
                     unsigned long long *__restrict__ next_string, uint32_t *__restrict__ redo_list,
                     unsigned long long *__restrict__ redo_n, const uint32_t *__restrict__ gate) {
    RXM_DYN_SMEM(smem);
    if (gate && *gate != 0u) return;  // a batch of long strings: K3 runs it (mfa_pick_kernel decided on the device)
    constexpr uint32_t ALL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    // ---- block-shared program tables ----
    uint32_t *s_begin = reinterpret_cast<uint32_t *>(smem);
    uint32_t *s_count = s_begin + n_keys;
    uint32_t *s_lbeg = s_count + n_keys;
    uint32_t *s_lcnt = s_lbeg + n_keys;
    uint16_t *s_sel = reinterpret_cast<uint16_t *>(s_lcnt + n_keys);
    size_t o = (size_t(n_keys) * 16 + size_t(n_sel) * 2 + 15) & ~size_t(15);
    ProgItem *s_items = reinterpret_cast<ProgItem *>(smem + o);
    if (items_in_smem) o += size_t(n_items) * sizeof(ProgItem);
    for (uint32_t k = threadIdx.x; k < n_keys; k += blockDim.x) {
        s_begin[k] = gp.begin[k];
        s_count[k] = gp.count[k];
        s_lbeg[k] = gp.lbeg[k];
        s_lcnt[k] = gp.lcnt[k];
    }
    for (uint32_t k = threadIdx.x; k < n_sel; k += blockDim.x) s_sel[k] = gp.sel[k];
    if (items_in_smem) {
        const uint4 *src = reinterpret_cast<const uint4 *>(gp.items);
        uint4 *dst = reinterpret_cast<uint4 *>(s_items);
        for (uint32_t k = threadIdx.x; k < n_items; k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();
    const K4Prog p{items_in_smem ? s_items : gp.items, s_begin, s_count, s_lbeg, s_lcnt, s_sel, gp.n_cells};

    K4Sim<NC, K4_THREADS> sim;
    sim.base = reinterpret_cast<uint32_t *>(smem + o) + threadIdx.x;
    sim.pool = maxl;
    sim.use_map = v.n_states <= K4_MAP_STATES && maxl <= 15u;
    sim.want = K4_WANT_NONE;

    // tickets: with the tile sort's records, ticket t is record (t mod 32) of group g of tile tl, group g
    // of every tile before group g+1 of any (the long strings of the whole batch first)
    const uint64_t ntiles = (n + K1_TILE_STRINGS - 1) / K1_TILE_STRINGS;
    const uint64_t tickets = recs ? ntiles * K1_TILE_STRINGS : n;
    unsigned long long wnext = 0, wend = 0;  // this warp's tickets in hand (warp-uniform)
    bool have = false, exhausted = false;
    uint64_t si = 0;
    for (;;) {
        const bool want = !have && !exhausted;
        const uint32_t wm = __ballot_sync(ALL, want);
        if (wm) {
            if (wnext == wend) {
                unsigned long long t0 = 0;
                if (lane == 0) t0 = atomicAdd(next_string, (unsigned long long)K4_CHUNK);
                t0 = __shfl_sync(ALL, t0, 0);
                wnext = t0;
                wend = t0 + K4_CHUNK;
            }
            const uint32_t rank = uint32_t(__popc(wm & ((1u << lane) - 1u)));
            const uint32_t avail = uint32_t(wend - wnext);
            if (want && rank < avail) {
                const uint64_t t = wnext + rank;
                if (t >= tickets) {
                    exhausted = true;
                } else {
                    bool skip = false;
                    si = t;
                    if (recs) {
                        const uint64_t g = t / (ntiles * 32u), rem = t - g * (ntiles * 32u);
                        const uint64_t tl = rem >> 5, pos = tl * K1_TILE_STRINGS + g * 32u + (rem & 31u);
                        if (pos >= min(n, (tl + 1u) * K1_TILE_STRINGS)) skip = true;  // the last tile is short
                        else si = recs[pos].idx;
                    }
                    if (!skip) {
                        const uint64_t sb = sp.begin[si], se = sp.end[si];
                        if (se - sb >= 0x7fffffffull) {
                            atomicAdd(overflow, 1ull);
                            out[si] = 0;
                        } else {
                            sim.start(chars + sb, uint32_t(se - sb), v.reversed, v.start);
                            have = true;
                        }
                    }
                }
            }
            const uint32_t took = uint32_t(__popc(wm));
            wnext += took < avail ? took : avail;
        }
        if (__all_sync(ALL, exhausted && !have)) break;
        // One round.  Every string says what it needs next -- a burst of repeated steps (phase A: a flat loop,
        // every lane in it runs the same few instructions) or one step in full (phase B) -- and the warp runs
        // the phase MOST of its strings want; the others wait for a round of their kind.  Lanes that wait
        // cost no issue slots, and each phase runs with most lanes of the warp in it.
        if (have) sim.pre();
        const uint32_t w = have ? sim.want : uint32_t(K4_WANT_NONE);
        const uint32_t mA = __ballot_sync(ALL, w == K4_WANT_A), mB = __ballot_sync(ALL, w == K4_WANT_B);
        if (mA != 0u && __popc(mA) >= __popc(mB)) {
            if (w == K4_WANT_A) sim.phase_a();
        } else if (mB != 0u) {
            if (w == K4_WANT_B) sim.phase_b(v, p);
        }
        __syncwarp(ALL);
        if (have && sim.want == K4_DONE) {
            int r = sim.result;
            if (r == 2) {
                r = 0;
                if (redo_list) redo_list[atomicAdd(redo_n, 1ull)] = uint32_t(si);
                else atomicAdd(overflow, 1ull);
            }
            out[si] = uint8_t(r);
            have = false;
        }
    }
}

template <int NC>
int launch_k4(const MfaView &v, const K4Prog &gp, uint32_t n_items, uint32_t n_keys, uint32_t n_sel, uint32_t maxl,
              const uint8_t *d_chars, Spans spans, const K1Rec *d_recs, uint64_t n, uint8_t *d_out,
              unsigned long long *d_overflow, unsigned long long *d_next, uint32_t *d_redo_list,
              unsigned long long *d_redo_n, int sm_count, uint32_t sharing, cudaStream_t stream, const uint32_t *d_gate) {
    const size_t tab = (size_t(n_keys) * 16 + size_t(n_sel) * 2 + 15) & ~size_t(15);
    const bool in_smem = size_t(n_items) * sizeof(ProgItem) <= 24 * 1024;
    const size_t smem = tab + (in_smem ? size_t(n_items) * sizeof(ProgItem) : 0) +
                        size_t(K4_THREADS) * k4_words(NC, maxl) * 4;
    if (smem > 200 * 1024) return RXM_ERR_UNSUPPORTED;
    auto kern = k4_mfa_thread_kernel<NC>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
        return RXM_ERR_CUDA;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, K4_THREADS, smem) != cudaSuccess || nb <= 0)
        return RXM_ERR_CUDA;
    uint64_t blocks = uint64_t(sm_count) * nb / (sharing ? sharing : 1u);
    if (blocks == 0) blocks = 1;
    const uint64_t need = (n + K4_THREADS - 1) / K4_THREADS;
    if (blocks > need) blocks = need;
    if (cudaMemsetAsync(d_next, 0, sizeof(unsigned long long), stream) != cudaSuccess) return RXM_ERR_CUDA;
    RXM_LAUNCH(kern, unsigned(blocks), K4_THREADS, smem, stream, v, gp, n_items, n_keys, n_sel, in_smem ? 1u : 0u, maxl,
               d_chars, spans, d_recs, n, d_out, d_overflow, d_next, d_redo_list, d_redo_n, d_gate);
    return RXM_OK;
}

}  // namespace

int k4_launch(const MfaView &v, const K4Prog &gp, uint32_t n_items, uint32_t n_keys, uint32_t n_sel, uint32_t n_cells,
              uint32_t maxl, const uint8_t *d_chars, Spans spans, const K1Rec *d_recs, uint64_t n, uint8_t *d_out,
              unsigned long long *d_overflow, unsigned long long *d_next, uint32_t *d_redo_list,
              unsigned long long *d_redo_n, int sm_count, uint32_t sharing, cudaStream_t stream, int *launched,
              const uint32_t *d_gate) {
    *launched = 0;
    int st;
    if (n_cells <= 1) st = launch_k4<1>(v, gp, n_items, n_keys, n_sel, maxl, d_chars, spans, d_recs, n, d_out, d_overflow, d_next, d_redo_list, d_redo_n, sm_count, sharing, stream, d_gate);
    else if (n_cells <= 2) st = launch_k4<2>(v, gp, n_items, n_keys, n_sel, maxl, d_chars, spans, d_recs, n, d_out, d_overflow, d_next, d_redo_list, d_redo_n, sm_count, sharing, stream, d_gate);
    else if (n_cells <= 4) st = launch_k4<4>(v, gp, n_items, n_keys, n_sel, maxl, d_chars, spans, d_recs, n, d_out, d_overflow, d_next, d_redo_list, d_redo_n, sm_count, sharing, stream, d_gate);
    else return RXM_ERR_UNSUPPORTED;
    if (st == RXM_OK) *launched = 1;
    return st;
}

}  // namespace rxm
