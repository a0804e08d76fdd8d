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

// Phase A's verification (rxm_k4_core.cuh) done by the WHOLE WARP for one string: the first j in [vp, vcap) with
// s[j] != s[j - delta], or vcap.  A thread that checks its own string reads 32 bytes per round trip from a
// line no other lane touches (ncu on config 3: 5.3 x the input over L2 -> SM, long-scoreboard stalls on every
// issue); here a round trip is 1 KB of consecutive addresses -- 32 lanes x one aligned 16-byte vector x 2 in
// flight -- and the vectors delta bytes back come from lines this warp has just read.  The stream delta bytes back
// stands at a fixed byte offset against the 16-byte grid, so its five words per vector are picked with a
// warp-uniform, loop-invariant word offset (WO: a template parameter, chosen once) and one funnel shift each.
// Only aligned 16-byte vectors that hold a byte of the string are loaded.  Warp-uniform arguments and result;
// needs vp >= delta.
constexpr uint32_t K4_COOP_MIN = 128;  // bytes still to verify from which a string can be worth the warp's time
constexpr int K4_COOP_UNROLL = 2;

__device__ __forceinline__ uint4 k4_ld128(const uint8_t *p) {  // p is 16-byte aligned
#if defined(__CUDA_ARCH__)
    return __ldg(reinterpret_cast<const uint4 *>(p));
#else
    return *reinterpret_cast<const uint4 *>(p);
#endif
}
// the bytes [lo, hi) of a 16-byte vector that fall into its word k, as a byte mask of that word
__device__ __forceinline__ uint32_t k4_word_mask(uint32_t lo, uint32_t hi, uint32_t k) {
    const int a = int(lo) - int(4u * k), z = int(hi) - int(4u * k);
    if (a >= 4 || z <= 0) return 0u;
    uint32_t m = 0xffffffffu;
    if (a > 0) m &= 0xffffffffu << (8 * a);
    if (z < 4) m &= (1u << (8 * z)) - 1u;
    return m;
}

template <int I>
__device__ __forceinline__ uint32_t k4_pick(const uint4 &a0, const uint4 &a1) {  // word I of the eight
    if constexpr (I == 0) return a0.x;
    else if constexpr (I == 1) return a0.y;
    else if constexpr (I == 2) return a0.z;
    else if constexpr (I == 3) return a0.w;
    else if constexpr (I == 4) return a1.x;
    else if constexpr (I == 5) return a1.y;
    else if constexpr (I == 6) return a1.z;
    else return a1.w;  // (I == 8 is only asked for when the streams are word-aligned: never shifted in)
}

template <int WO>
__device__ __forceinline__ uint32_t k4_coop_loop(const uint8_t *s, const uint8_t *end, const uint8_t *b, const uint8_t *lim,
                                                 const uint8_t *base0, uint32_t delta, uint32_t sh16, uint32_t lane) {
    const uint32_t bo8 = (sh16 & 3u) * 8u;
    constexpr uint32_t NONE = 0xffffffffu;
    for (const uint8_t *base = base0; base < lim; base += 512 * K4_COOP_UNROLL) {
        uint32_t at[K4_COOP_UNROLL];  // where this lane's vector differs first (offset into the string), or NONE
RXM_UNROLL
        for (int u = 0; u < K4_COOP_UNROLL; u++) {
            const uint8_t *c = base + 512 * u + 16u * lane;  // this lane's vector of the stream
            at[u] = NONE;
            if (c < lim) {
                const uint4 w = k4_ld128(c);
                const uint8_t *A = c - delta - sh16;  // aligned: the two vectors that hold [c - delta, c - delta + 16)
                const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                const uint4 a0 = (A + 16 > s) ? k4_ld128(A) : z;
                const uint4 a1 = (sh16 != 0u && A + 16 < end) ? k4_ld128(A + 16) : z;
                const uint32_t v0 = k4_pick<WO>(a0, a1), v1 = k4_pick<WO + 1>(a0, a1), v2 = k4_pick<WO + 2>(a0, a1),
                               v3 = k4_pick<WO + 3>(a0, a1), v4 = k4_pick<WO + 4>(a0, a1);
                uint32_t d0 = (bo8 ? __funnelshift_r(v0, v1, bo8) : v0) ^ w.x;
                uint32_t d1 = (bo8 ? __funnelshift_r(v1, v2, bo8) : v1) ^ w.y;
                uint32_t d2 = (bo8 ? __funnelshift_r(v2, v3, bo8) : v2) ^ w.z;
                uint32_t d3 = (bo8 ? __funnelshift_r(v3, v4, bo8) : v3) ^ w.w;
                if (c < b || c + 16 > lim) {  // the vectors that hold vp and vcap: only the bytes in [vp, vcap) count
                    const uint32_t lo = c < b ? uint32_t(b - c) : 0u, hi = c + 16 > lim ? uint32_t(lim - c) : 16u;
                    d0 &= k4_word_mask(lo, hi, 0u);
                    d1 &= k4_word_mask(lo, hi, 1u);
                    d2 &= k4_word_mask(lo, hi, 2u);
                    d3 &= k4_word_mask(lo, hi, 3u);
                }
                if (d0 | d1 | d2 | d3) {  // (rare: at most once per string and round)
                    const uint32_t k = d0 ? 0u : (d1 ? 1u : (d2 ? 2u : 3u));
                    const uint32_t dk = d0 ? d0 : (d1 ? d1 : (d2 ? d2 : d3));
                    at[u] = uint32_t(c - s) + 4u * k + (uint32_t(k4_ffs(dk)) - 1u) / 8u;
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
    return uint32_t(lim - s);
}

__device__ __forceinline__ uint32_t k4_coop_verify(const uint8_t *s, uint32_t n, uint32_t delta, uint32_t vp, uint32_t vcap,
                                                   uint32_t lane) {
    const uint8_t *end = s + n, *b = s + vp, *lim = s + vcap;
    const uint8_t *base0 = b - (reinterpret_cast<uintptr_t>(b) & 15u);
    const uint32_t sh16 = uint32_t(reinterpret_cast<uintptr_t>(base0) - delta) & 15u;
    switch (sh16 >> 2) {
        case 0: return k4_coop_loop<0>(s, end, b, lim, base0, delta, sh16, lane);
        case 1: return k4_coop_loop<1>(s, end, b, lim, base0, delta, sh16, lane);
        case 2: return k4_coop_loop<2>(s, end, b, lim, base0, delta, sh16, lane);
        default: return k4_coop_loop<3>(s, end, b, lim, base0, delta, sh16, lane);
    }
}

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
        // strings with a long stretch still to verify: the warp checks them one after another, all lanes on one
        // string (k4_coop_verify); their own phase A then only moves the answered steps on
        // (worth it from ~24 bytes per string that wants phase A: their own loop checks all of them at once, 32 bytes
        // per ~150 instructions, the warp one after another, 1 KB per ~70 and ~60 to set up)
        const uint32_t coop_min = max(K4_COOP_MIN, 24u * uint32_t(__popc(__ballot_sync(ALL, w == K4_WANT_A))));
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
