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
// shared-memory bytes of the block-shared program tables: the packed lists (K4Prog::lists), sel [n_sel]; rounded up to 16
RXM_HD size_t k4_table_bytes(uint32_t n_keys, uint32_t n_sel, uint32_t n_classes) {
    return (size_t(k4_list_words(n_keys, n_classes)) * 4 + size_t(n_sel) * 2 + 15) & ~size_t(15);
}
constexpr uint32_t K4_CHUNK = 64;  // tickets a warp takes from the global counter at a time
#ifndef RXM_K4_WAIT_MAX  // (tuning builds set it)
#define RXM_K4_WAIT_MAX 3
#endif
constexpr uint32_t K4_WAIT_MAX = RXM_K4_WAIT_MAX;  // rounds a string may wait for its phase

// Phase A's verification (rxm_k4_core.cuh) done by the WHOLE WARP for one string: the first j in [vp, vcap) with
// s[j] != s[j - delta], or vcap.  A thread that checks its own string reads 32 bytes per round trip from a
// line no other lane touches (ncu on config 3: 5.3 x the input over L2 -> SM, long-scoreboard stalls on every
// issue); here a round trip is 512 bytes of consecutive addresses -- 32 lanes x one aligned 16-byte vector --
// and the vectors delta bytes back come from lines this warp has just read.  The stream delta bytes back
// stands at a fixed byte offset against the 16-byte grid, so its five words per vector are picked with a
// warp-uniform, loop-invariant word offset (selects on three uniform predicates: four copies of the loop, one per
// offset, cost more in instruction-cache misses than they saved -- ncu: no_instruction 10.7 per issue) and one
// funnel shift each.
// Only aligned 16-byte vectors that hold a byte of the string are loaded.  Warp-uniform arguments and result;
// needs vp >= delta.
#ifndef RXM_K4_COOP_MIN  // (tuning builds set these two)
#define RXM_K4_COOP_MIN 512
#endif
#ifndef RXM_K4_COOP_PER_LANE
#define RXM_K4_COOP_PER_LANE 48
#endif
constexpr uint32_t K4_COOP_MIN = RXM_K4_COOP_MIN;  // bytes still to verify from which a string can be worth the warp's time
#ifndef RXM_K4_COOP_UNROLL  // (tuning builds set it)
#define RXM_K4_COOP_UNROLL 1  // measured on config 3: 1 -> 2.55 ms, 2 -> 2.74, 4 -> 3.26 (vectors past the stretch's end cost more than a round trip saves)
#endif
constexpr int K4_COOP_UNROLL = RXM_K4_COOP_UNROLL;

__device__ __forceinline__ uint4 k4_ld128(const uint8_t *p) {  // p is 16-byte aligned
#if defined(__CUDA_ARCH__)
    return __ldg(reinterpret_cast<const uint4 *>(p));
#else
#ifdef RXM_SIMT_HOST  // the emulator reports what the device would fault on
    if (reinterpret_cast<uintptr_t>(p) & 15u) simt::fail(4, "misaligned 16-byte global load");
#endif
    return *reinterpret_cast<const uint4 *>(p);
#endif
}
// the bytes [lo, hi) of a 16-byte vector that fall into its word k, as a byte mask of that word (branch-free)
__device__ __forceinline__ uint32_t k4_word_mask(uint32_t lo, uint32_t hi, uint32_t k) {
    const int a = min(max(int(lo) - int(4u * k), 0), 4), z = min(max(int(hi) - int(4u * k), 0), 4);
    return uint32_t(~0ull << (8 * a)) & uint32_t((1ull << (8 * z)) - 1ull);
}

__device__ __forceinline__ uint32_t k4_coop_verify(const uint8_t *s, uint32_t n, uint32_t delta, uint32_t vp, uint32_t vcap,
                                                   uint32_t lane) {
    // 32-bit offsets from an aligned origin 16..31 bytes before the string (so that nothing below goes negative)
    const uint32_t lo_s = 16u + uint32_t(reinterpret_cast<uintptr_t>(s) & 15u);  // the string is [lo_s, end_s)
    const uint8_t *org = s - lo_s;
    const uint32_t end_s = lo_s + n, b = lo_s + vp, lim = lo_s + vcap;
    const uint32_t sh16 = (0u - delta) & 15u;  // (any 16-aligned offset - delta) & 15: where the stream delta back stands
    const uint32_t bo8 = (sh16 & 3u) * 8u;
    const bool w1 = (sh16 >> 2) == 1u, w2 = (sh16 >> 2) == 2u, w3 = (sh16 >> 2) == 3u;  // (uniform: selects, not branches)
    constexpr uint32_t NONE = 0xffffffffu;
    for (uint32_t base = b & ~15u; base < lim; base += 512u * K4_COOP_UNROLL) {  // (warp-uniform)
        uint32_t at[K4_COOP_UNROLL];  // where this lane's vector differs first (offset into the string), or NONE
RXM_UNROLL
        for (int u = 0; u < K4_COOP_UNROLL; u++) {
            const uint32_t c = base + 512u * uint32_t(u) + 16u * lane;  // this lane's vector of the stream
            at[u] = NONE;
            if (c < lim) {
                const uint4 w = k4_ld128(org + c);
                const uint32_t A = c - delta - sh16;  // aligned: the two vectors that hold [c - delta, c - delta + 16)
                const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                const uint4 a0 = (A + 16u > lo_s) ? k4_ld128(org + A) : z;
                const uint4 a1 = (sh16 != 0u && A + 16u < end_s) ? k4_ld128(org + A + 16u) : z;
                const uint32_t v0 = w3 ? a0.w : (w2 ? a0.z : (w1 ? a0.y : a0.x));
                const uint32_t v1 = w3 ? a1.x : (w2 ? a0.w : (w1 ? a0.z : a0.y));
                const uint32_t v2 = w3 ? a1.y : (w2 ? a1.x : (w1 ? a0.w : a0.z));
                const uint32_t v3 = w3 ? a1.z : (w2 ? a1.y : (w1 ? a1.x : a0.w));
                const uint32_t v4 = w3 ? a1.w : (w2 ? a1.z : (w1 ? a1.y : a1.x));
                uint32_t d0 = (bo8 ? __funnelshift_r(v0, v1, bo8) : v0) ^ w.x;
                uint32_t d1 = (bo8 ? __funnelshift_r(v1, v2, bo8) : v1) ^ w.y;
                uint32_t d2 = (bo8 ? __funnelshift_r(v2, v3, bo8) : v2) ^ w.z;
                uint32_t d3 = (bo8 ? __funnelshift_r(v3, v4, bo8) : v3) ^ w.w;
                if (c < b || c + 16u > lim) {  // the vectors that hold vp and vcap: only the bytes in [vp, vcap) count
                    const uint32_t lo = c < b ? b - c : 0u, hi = c + 16u > lim ? lim - c : 16u;
                    d0 &= k4_word_mask(lo, hi, 0u);
                    d1 &= k4_word_mask(lo, hi, 1u);
                    d2 &= k4_word_mask(lo, hi, 2u);
                    d3 &= k4_word_mask(lo, hi, 3u);
                }
                if (d0 | d1 | d2 | d3) {  // (rare: at most once per string and round)
                    const uint32_t k = d0 ? 0u : (d1 ? 1u : (d2 ? 2u : 3u));
                    const uint32_t dk = d0 ? d0 : (d1 ? d1 : (d2 ? d2 : d3));
                    at[u] = c - lo_s + 4u * k + (uint32_t(k4_ffs(dk)) - 1u) / 8u;
                }
            }
        }
        uint32_t first = NONE;  // the earliest difference of the round: vectors are in stream order by (u, lane)
RXM_UNROLL
        for (int u = K4_COOP_UNROLL - 1; u >= 0; u--) {
            const uint32_t bad = __ballot_sync(0xffffffffu, at[u] != NONE);
            if (bad) first = __shfl_sync(0xffffffffu, at[u], k4_ffs(bad) - 1);
        }
        if (first != NONE) return first;
    }
    return vcap;
}

template <int NC>
__global__ void __launch_bounds__(K4_THREADS, 6)
k4_mfa_thread_kernel(MfaView v, K4Prog gp, uint32_t n_items, uint32_t n_keys, uint32_t n_sel, uint32_t items_in_smem,
                     uint32_t maxl, const uint8_t *__restrict__ chars, const Spans sp, const K1Rec *__restrict__ recs,
                     uint64_t n, uint8_t *__restrict__ out, unsigned long long *__restrict__ overflow,
                     unsigned long long *__restrict__ next_string, uint32_t *__restrict__ redo_list,
                     unsigned long long *__restrict__ redo_n, const uint32_t *__restrict__ gate) {
    RXM_DYN_SMEM(smem);
    if (gate && *gate != 0u) return;  // a batch of long strings: K3 runs it (mfa_pick_kernel decided on the device)
    constexpr uint32_t ALL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    // ---- block-shared program tables: the packed lists, the item selections, the items if they fit ----
    const uint32_t n_lw = k4_list_words(n_keys, gp.n_classes);
    uint32_t *s_lists = reinterpret_cast<uint32_t *>(smem);
    uint16_t *s_sel = reinterpret_cast<uint16_t *>(s_lists + n_lw);
    size_t o = k4_table_bytes(n_keys, n_sel, gp.n_classes);
    ProgItem *s_items = reinterpret_cast<ProgItem *>(smem + o);
    if (items_in_smem) o += size_t(n_items) * sizeof(ProgItem);
    for (uint32_t k = threadIdx.x; k < n_lw; k += blockDim.x) s_lists[k] = gp.lists[k];
    for (uint32_t k = threadIdx.x; k < n_sel; k += blockDim.x) s_sel[k] = gp.sel[k];
    if (items_in_smem) {
        const uint4 *src = reinterpret_cast<const uint4 *>(gp.items);
        uint4 *dst = reinterpret_cast<uint4 *>(s_items);
        for (uint32_t k = threadIdx.x; k < n_items; k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();
    const K4Prog p{items_in_smem ? s_items : gp.items, s_lists, s_sel, gp.n_cells, gp.n_classes, n_keys};

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
    uint32_t waited = 0;  // rounds this lane's string has waited for its phase
#ifndef RXM_K4_HIST_MIN
#define RXM_K4_HIST_MIN 24
#endif
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
        // strings with a long stretch still to verify: the warp checks them one after another, all lanes on one
        // string (k4_coop_verify); their own phase A then only moves the answered steps on
        // (worth it from 512 bytes, and from 48 bytes per string that wants phase A: their own loop checks all of them
        // at once, 32 bytes per ~150 instructions; the warp takes them one after another.  Measured on the ten README
        // examples' short strings: lower thresholds cost examples 6, 8 and 10 up to 14 %, this one nothing.)
        const uint32_t coop_min = max(K4_COOP_MIN, uint32_t(RXM_K4_COOP_PER_LANE) * uint32_t(__popc(__ballot_sync(ALL, w == K4_WANT_A))));
        for (uint32_t mC = __ballot_sync(ALL, w == K4_WANT_A && !sim.rp_mism && sim.rp_vp + coop_min <= sim.rp_vcap); mC;
             mC &= mC - 1u) {
            const int l = k4_ffs(mC) - 1;
            const uint8_t *cs = reinterpret_cast<const uint8_t *>(
                uintptr_t(__shfl_sync(ALL, uint64_t(reinterpret_cast<uintptr_t>(sim.s)), l)));
            const uint32_t cn = __shfl_sync(ALL, sim.n, l), cd = __shfl_sync(ALL, sim.rp_delta, l);
            const uint32_t cvp = __shfl_sync(ALL, sim.rp_vp, l), cvc = __shfl_sync(ALL, sim.rp_vcap, l);
            const uint32_t pos = k4_coop_verify(cs, cn, cd, cvp, cvc, lane);
            if (int(lane) == l) {
                sim.rp_vp = pos;
                sim.rp_mism = pos < cvc;
            }
        }
        const uint32_t mA = __ballot_sync(ALL, w == K4_WANT_A), mB = __ballot_sync(ALL, w == K4_WANT_B);
        // the phase most strings want -- but a string never waits more than K4_WAIT_MAX rounds for its phase: then the
        // other phase runs once, whatever the count.  (The plain majority starves the strings of the rarer phase until
        // enough of them have piled up: example 9 of the README runs mostly full steps, the few strings at the start of
        // an a-run sat out ~a third of the warp's rounds -- 21.8 -> 15.5 ms; config 3 is unchanged, 2.60 ms.  Serving the
        // rarer phase at once, or both phases every round, costs config 3 and example 10 12 - 40 %: `tools/k4_ab.sh`.)
        bool run_a = mA != 0u && __popc(mA) >= __popc(mB);
        if (mA != 0u && mB != 0u) {
            const bool starving = have && waited >= K4_WAIT_MAX && (w == K4_WANT_A) != run_a;
            if (__any_sync(ALL, starving)) run_a = !run_a;
        }
        if (run_a) {
            if (w == K4_WANT_A) sim.phase_a();
        } else if (mB != 0u) {
            if (w == K4_WANT_B) sim.phase_b(v, p);
        }
        waited = (have && ((w == K4_WANT_A) != run_a)) ? waited + 1u : 0u;
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
    const size_t tab = k4_table_bytes(n_keys, n_sel, gp.n_classes);
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
