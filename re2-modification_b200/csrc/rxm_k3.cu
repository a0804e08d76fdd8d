// K3 -- MFA, one WARP per string, sm_100a.
// Replaces MFA::match (mfa.cpp:215-236) for whole batches; meant for long strings and
// large automata, where one thread per string (K2) serialises everything.
//
// The warp keeps the successor set as one slot per node in shared memory.
// A slot is a 64-bit ORDER KEY  first:28 | lowest cell name:4 | creation stamp:32  -- the
// reference's std::set<MemoryState> order for one node -- plus the cell payload.  In a step
// ALL live configurations are expanded at once: the items of their host-compiled edge
// programs (rxm_plan.cpp: compile_programs) are laid end to end and dealt to the lanes, 32
// per pass; a lane builds the successor its item produces (a call entry or a non-recursive
// edge) and claims the target slot with atomicMin on the key; the lane whose key stands
// after the warp has synchronised writes the payload.  Creation order needs no counter:
// prog_stamp (rxm_mfa_core.cuh) orders exactly where the reference's order is consulted.
// Backreference blocks are compared by the whole warp, 128 bytes per iteration.
// Idle steps (every configuration waiting inside a block) are skipped exactly as in K2.
#include "rxm_kernels.cuh"

namespace rxm {

namespace {

constexpr uint64_t K3_EMPTY = ~0ull;
#ifndef RXM_K3_MIN_BLOCKS
#define RXM_K3_MIN_BLOCKS 4  // blocks per SM the several-strings-per-warp kernels are compiled for: 64 registers, no spills
                             // (config 3: 43.6 ms against 46.4 at 5 blocks / 48 registers and 44.6 at 3 / 80)
#endif
constexpr int K3_WARPS = 8;

__device__ __forceinline__ uint64_t k3_key(uint32_t first, uint32_t flags, uint32_t born) {
    return (uint64_t(first) << 36) | (uint64_t(lowvar(flags)) << 32) | born;
}

// Collectives of a tile.  The tiles of a warp run in lock-step and ALL control flow around a
// collective is warp-uniform (trip counts are the maximum over the warp's tiles, the bodies are
// predicated), so every vote / shuffle is issued once with the full mask and a tile takes its own
// bits.  With per-tile masks the compiler has to group the lanes by mask value first (MATCH.ANY +
// one VOTE / ENDCOLLECTIVE round per tile): measured at 28 % of the stall samples of the kernel.
template <int TILE>
struct TileOps {
    static constexpr uint32_t ALL = 0xffffffffu;
    static constexpr uint32_t TM = (TILE == 32) ? 0xffffffffu : ((1u << (TILE & 31)) - 1u);
    uint32_t tshift;  // first warp lane of this tile
    __device__ __forceinline__ uint32_t ballot(bool p) const { return (__ballot_sync(ALL, p) >> tshift) & TM; }
    __device__ __forceinline__ bool any(bool p) const { return ballot(p) != 0u; }
    __device__ __forceinline__ bool all(bool p) const { return ballot(p) == TM; }
    // p has the same value in every lane of a tile: is it set in any / every tile of the warp?
    // (TILE == 32: the tile is the warp, no vote)
    __device__ __forceinline__ static bool warp_any(bool p) { return TILE == 32 ? p : __any_sync(ALL, p); }
    __device__ __forceinline__ static bool warp_all(bool p) { return TILE == 32 ? p : __all_sync(ALL, p); }
    __device__ __forceinline__ uint32_t min(uint32_t x) const {
        if (TILE == 32) return __reduce_min_sync(ALL, x);
#pragma unroll
        for (int d = TILE / 2; d >= 1; d >>= 1) {
            const uint32_t y = __shfl_xor_sync(ALL, x, d);
            x = y < x ? y : x;
        }
        return x;
    }
};

// mem[pa, pa+L) == mem[pb, pb+L) ?  per tile, 4 bytes per lane per iteration, any alignment.
// Called by the whole warp; a tile without a comparison passes L == 0.
template <int TILE>
__device__ __forceinline__ bool tile_span_equal(const TileOps<TILE> &to, const uint8_t *pa, const uint8_t *pb, uint32_t L,
                                                uint32_t lane) {
    bool eq = true;
    bool live = !(pa == pb || L == 0);
    const uint32_t ba = uint32_t(reinterpret_cast<uintptr_t>(pa)) & 3u, bb = uint32_t(reinterpret_cast<uintptr_t>(pb)) & 3u;
    const uint32_t *wa = reinterpret_cast<const uint32_t *>(pa - ba);
    const uint32_t *wb = reinterpret_cast<const uint32_t *>(pb - bb);
    for (uint32_t base = 0;; base += 4u * TILE) {
        live = live && base < L;
        if (!TileOps<TILE>::warp_any(live)) break;
        const uint32_t off = base + lane * 4u;
        bool ne = false;
        if (live && off < L) {
            const uint32_t rem = L - off, take = rem < 4u ? rem : 4u;
            const uint32_t idx = off >> 2;
            const uint32_t a0 = __ldg(wa + idx), b0 = __ldg(wb + idx);
            const uint32_t a1 = (ba + take > 4u) ? __ldg(wa + idx + 1) : 0u;  // only if bytes < L live there
            const uint32_t b1 = (bb + take > 4u) ? __ldg(wb + idx + 1) : 0u;
            uint32_t x = __funnelshift_r(a0, a1, ba * 8u) ^ __funnelshift_r(b0, b1, bb * 8u);
            if (take < 4u) x &= (1u << (8u * take)) - 1u;
            ne = x != 0u;
        }
        if (to.any(ne)) {
            eq = false;
            live = false;
        }
    }
    return eq;
}

template <int NC>
__device__ __forceinline__ uint32_t exists_mask_n(uint32_t flags) {  // bit k <- flags bit 3k, k < NC
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < NC; k++) m |= ((flags >> (3 * k)) & 1u) << k;
    return m;
}

template <int NC>
struct K3Cfg {  // a configuration in registers
    uint32_t first, born, flags, node;
    uint32_t start[NC], len[NC];
};

template <int NC>
__device__ __forceinline__ void k3_working(K3Cfg<NC> &w, const K3Cfg<NC> &root, uint32_t created,
                                           uint32_t created_open, uint32_t marks) {
    w = root;
#pragma unroll
    for (int k = 0; k < NC; k++) {
        if ((created >> k) & 1u) {
            const uint32_t fl = ((created_open >> k) & 1u) ? 3u : 1u;
            w.flags = (w.flags & ~(7u << (3 * k))) | (fl << (3 * k));
            w.start[k] = root.first;
            w.len[k] = 0;
        }
        if ((marks >> k) & 1u) w.flags |= 4u << (3 * k);
    }
}

// doMemoryWriteActions (mfa.cpp:80-105) -- same as apply_actions<NC> on K3Cfg
template <int NC>
__device__ __forceinline__ void k3_apply(K3Cfg<NC> &m, uint32_t open_mask, uint32_t close_mask, uint32_t tstart,
                                         uint32_t tlen) {
#pragma unroll
    for (int k = 0; k < NC; k++) {
        const uint32_t o = (open_mask >> k) & 1u, c = (close_mask >> k) & 1u;
        uint32_t f = (m.flags >> (3 * k)) & 7u;
        if (o) {
            f = 3u;
            m.start[k] = tstart;
            m.len[k] = tlen;
        } else if (f & 1u) {
            if (c) f &= ~2u;
            else if (f & 2u) {
                if (m.len[k] == 0) m.start[k] = tstart;
                m.len[k] += tlen;
            }
        }
        m.flags = (m.flags & ~(7u << (3 * k))) | (f << (3 * k));
    }
}

template <int NC>
__device__ __forceinline__ uint32_t k3_need(uint32_t flags, const uint32_t *len) {  // is_siffix_long_enough
    uint32_t need = 0;
#pragma unroll
    for (int k = 0; k < NC; k++) {
        const uint32_t fl = (flags >> (3 * k)) & 7u;
        if ((fl & 1u) && ((fl & 2u) || !(fl & 4u))) need += len[k];
    }
    return need;
}

// TILE lanes cooperate on one string (TILE = 32: the whole warp; 16 / 8: two / four strings
// per warp when the edge programs are short -- the items of a step fit one pass anyway).
// Several strings per warp (TILE < 32): throughput work, issue-bound -- 5 blocks per SM (48
// registers, a few spilled words) measured 17 % faster than 3 blocks at 73 registers.  TILE == 32
// is chosen for long strings, where the single longest string bounds the batch: full registers.
template <int NC, int TILE>
__global__ void __launch_bounds__(K3_WARPS * 32, TILE == 32 ? 1 : RXM_K3_MIN_BLOCKS)
k3_mfa_warp_kernel(MfaView v, ProgView gp, uint32_t n_items, uint32_t n_keys, uint32_t items_in_smem,
                   const uint8_t *__restrict__ chars, const Spans sp, const K1Rec *__restrict__ recs, uint64_t n,
                   uint8_t *__restrict__ out, unsigned long long *__restrict__ overflow,
                   unsigned long long *__restrict__ next_string, const uint32_t *__restrict__ list,
                   const unsigned long long *__restrict__ list_n, const uint32_t *__restrict__ gate, uint32_t gate_want) {
    RXM_DYN_SMEM(smem);
    // gate != null: this launch only runs if the batch is (gate_want = 1) / is not (0) one of long strings -- decided
    // on the device (mfa_pick_kernel), so that the call stays asynchronous when only the device knows the lengths
    if (gate && (*gate != 0u) != (gate_want != 0u)) return;
    // list != null: the strings to run are list[0 .. *list_n) (K4 hands over the strings that outgrew its
    // per-thread sets; the count is only known on the device)
    if (list) n = *list_n;
    constexpr uint32_t ALL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & (TILE - 1u);              // rank inside the tile
    const uint32_t tile = threadIdx.x / TILE;                     // tile index inside the block
    TileOps<TILE> to;
    to.tshift = (threadIdx.x & 31u) & ~(TILE - 1u);
    // ---- block-shared program tables ----
    uint32_t *s_begin = reinterpret_cast<uint32_t *>(smem);
    uint32_t *s_count = s_begin + n_keys;
    size_t o = (size_t(n_keys) * 8 + 15) & ~size_t(15);
    ProgItem *s_items = reinterpret_cast<ProgItem *>(smem + o);
    if (items_in_smem) o += size_t(n_items) * sizeof(ProgItem);
    for (uint32_t i = threadIdx.x; i < n_keys; i += blockDim.x) {
        s_begin[i] = gp.begin[i];
        s_count[i] = gp.count[i];
    }
    if (items_in_smem) {
        const uint4 *src = reinterpret_cast<const uint4 *>(gp.items);
        uint4 *dst = reinterpret_cast<uint4 *>(s_items);
        for (uint32_t i = threadIdx.x; i < n_items; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const ProgItem *items = items_in_smem ? s_items : gp.items;
    // ---- per-tile frontier: two buffers of SP slots ----
    const uint32_t SP = (v.n_states + TILE - 1u) & ~(TILE - 1u);  // slots, a whole number of tile passes
    const size_t per_tile = size_t(SP) * 2 * (8 + 4 + 8 * NC) + size_t(SP) * 12 + 16;
    uint8_t *wb = smem + o + size_t(tile) * per_tile;
    uint64_t *keys = reinterpret_cast<uint64_t *>(wb);                    // [2][SP]
    uint32_t *flg = reinterpret_cast<uint32_t *>(wb + size_t(SP) * 16);   // [2][SP]
    uint32_t *stt = flg + 2 * SP;                                         // [2][SP][NC]
    uint32_t *lnn = stt + 2 * SP * NC;                                    // [2][SP][NC]
    uint32_t *l_node = lnn + 2 * SP * NC;                                 // [SP]   live list: node,
    uint32_t *l_pb = l_node + SP;                                         // [SP]   first program item,
    uint32_t *l_off = l_pb + SP;                                          // [SP+1] start in the laid-out items

    // All tiles of a warp run the loop below in LOCK-STEP: one iteration is, per tile, "take the
    // next string" (when it has none) and then "one step of the string in hand".  Control flow is
    // warp-uniform -- a tile without work rides along predicated off -- so the tiles execute the
    // same instructions on different strings and every collective runs once for the whole warp.
    bool have_str = false, exhausted = false;
    unsigned long long si = 0;
    const uint8_t *s = nullptr;
    uint32_t n32 = 0, cur = 0, i = 0;
    bool ovf = false;
    for (;;) {
#ifdef RXM_SIMT_HOST
        if (threadIdx.x == 0) rxm_k3_simt_iterations++;  // tests/hostsim/kernels_simt.cpp: lock-step iterations of the warp
#endif
        const bool want = !have_str && !exhausted;
        if (TileOps<TILE>::warp_any(want)) {
            unsigned long long t = 0;
            if (want && lane == 0) t = atomicAdd(next_string, 1ull);
            t = __shfl_sync(ALL, t, 0, TILE);
            if (want) {  // no collective in here
                si = t;
                bool skip = false;
                if (recs) {
                    // strings are handed out in the tile sort's order -- group g (32 records) of every tile
                    // of 4096 before group g+1 of any: the long strings of the whole batch go first, so
                    // the tail of the launch is made of short ones
                    const uint64_t ntiles = (n + K1_TILE_STRINGS - 1) / K1_TILE_STRINGS;
                    const uint64_t g = si / (ntiles * 32u), rem = si - g * (ntiles * 32u);
                    const uint64_t tl = rem >> 5, pos = tl * K1_TILE_STRINGS + g * 32u + (rem & 31u);
                    if (g >= K1_TILE_STRINGS / 32u) si = n;                              // all handed out
                    else if (pos >= min(n, (tl + 1u) * K1_TILE_STRINGS)) skip = true;    // the last tile is short
                    else si = recs[pos].idx;
                }
                if (skip) {
                    // nothing this round; the next iteration takes another ticket
                } else if (si >= n) {
                    exhausted = true;
                } else {
                    if (list) si = list[si];
                    const uint64_t sb = sp.begin[si], se = sp.end[si];
                    if (se - sb >= (1ull << 28)) {  // first must fit 28 bits of the order key
                        if (lane == 0) {
                            atomicAdd(overflow, 1ull);
                            out[si] = 0;
                        }
                    } else {
                        s = chars + sb;
                        n32 = uint32_t(se - sb);
                        for (uint32_t q = lane; q < 2 * SP; q += TILE) {
                            const bool st0 = (q == v.start);  // (0, start, {})  mfa.cpp:217-219
                            keys[q] = st0 ? k3_key(0, 0, 0) : K3_EMPTY;
                            if (st0) flg[q] = 0;
                        }
                        cur = 0;
                        i = 0;
                        ovf = false;
                        have_str = true;
                    }
                }
            }
            __syncwarp(ALL);
        }
        if (TileOps<TILE>::warp_all(exhausted && !have_str)) break;
        const bool act = have_str;  // this tile has a string in hand
        bool finished = false;
        {
            uint64_t *K = keys + cur * SP, *KN = keys + (cur ^ 1u) * SP;
            // ---- A. list the live configurations and lay their programs end to end ----
            uint32_t m = 0, T = 0;
            bool any_active = false;
            for (uint32_t q = lane; q < SP; q += TILE) {  // warp-uniform trip count
                const uint64_t k = act ? K[q] : K3_EMPTY;
                if (act) KN[q] = K3_EMPTY;
                const bool lv = (k != K3_EMPTY);
                uint32_t pb = 0, pc = 0;
                if (lv) {
                    const uint32_t f = uint32_t(k >> 36);
                    const uint32_t fl = flg[cur * SP + q];
                    any_active |= (f == i);
                    bool pruned = false;
                    if (v.reversed && !(q == v.finish && f == n32)) {  // mfa.cpp:141
                        uint32_t need = 0;
#pragma unroll
                        for (int kk = 0; kk < NC; kk++) {
                            const uint32_t c3 = (fl >> (3 * kk)) & 7u;
                            if ((c3 & 1u) && ((c3 & 2u) || !(c3 & 4u))) need += lnn[(cur * SP + q) * NC + kk];
                        }
                        pruned = need > n32 - i;
                    }
                    if (!pruned) {
                        const uint32_t pkey = (q << gp.n_cells) | exists_mask_n<NC>(fl);
                        pb = s_begin[pkey];
                        if (pb == 0xffffffffu) ovf = true;  // a (node, cells) pair the host analysis missed
                        else pc = s_count[pkey] & kProgCountMask;
                    }
                }
                const uint32_t bal = to.ballot(lv);
                uint32_t inc = pc;  // inclusive scan of the item counts over the tile's lanes
#pragma unroll
                for (int d = 1; d < TILE; d <<= 1) {
                    const uint32_t u = __shfl_up_sync(ALL, inc, d, TILE);
                    if (int(lane) >= d) inc += u;
                }
                const uint32_t before = inc - pc, total = __shfl_sync(ALL, inc, TILE - 1, TILE);
                if (lv) {
                    const uint32_t j = m + __popc(bal & ((1u << lane) - 1u));
                    l_node[j] = q;
                    l_pb[j] = pb;
                    l_off[j] = T + before;
                }
                m += __popc(bal);
                T += total;
            }
            ovf = to.any(ovf);
            any_active = to.any(any_active);
            const bool dead = act && i < n32 && m == 0;  // :224-225
            const bool run = act && !dead;
            if (dead) finished = true;
            if (run && lane == 0) l_off[m] = T;
            __syncwarp(ALL);
            const uint32_t ch = (run && i < n32) ? uint32_t(v.reversed ? s[n32 - 1u - i] : s[i]) : 0u;
            const uint32_t digit_bit = (run && i < n32 && ch >= '1' && ch <= '9') ? (1u << (ch - '1')) : 0u;
            bool next_near = false;  // the new set holds a configuration that is active (or dead) at step i+1
            // what C0 needs to know about the new set, gathered from the lanes that write it: every
            // configuration stable and waiting, and the earliest event (activation / reversed-mode pruning)
            bool w_stable = true, w_any = false;
            uint32_t w_ev = n32;
            // ---- B. expand: one item per lane per pass ----
            const uint32_t Tmax = __reduce_max_sync(ALL, run ? T : 0u);
            for (uint32_t t0 = 0; t0 < Tmax; t0 += TILE) {
                const uint32_t tt = t0 + lane;
                bool have = false, need_cmp = false;
                uint32_t cmp_vs = 0, cmp_L = 0;
                K3Cfg<NC> cand, root;
                ProgItem it{};
                if (run && tt < T) {
                    uint32_t j = 0;
                    while (l_off[j + 1] <= tt) j++;
                    const uint32_t x = tt - l_off[j], rn = l_node[j];
                    const uint64_t rk = K[rn];
                    root.first = uint32_t(rk >> 36);
                    root.born = 0;
                    root.node = rn;
                    root.flags = flg[cur * SP + rn];
#pragma unroll
                    for (int k = 0; k < NC; k++) {
                        root.start[k] = stt[(cur * SP + rn) * NC + k];
                        root.len[k] = lnn[(cur * SP + rn) * NC + k];
                    }
                    const bool fin = (root.first == n32);
                    const bool active = (i != n32 && i == root.first);
                    const bool waiting = (i != n32 && i < root.first);
                    it = items[l_pb[j] + x];
                    if (!(fin && pi_skip_final(it))) {
                        if (!pi_is_leaf(it)) {
                            const uint32_t vv = pi_node(it);
                            if ((vv == v.finish && fin) || (waiting && pi_has_leaf(it))) {  // :138-140 / :195-197
                                k3_working<NC>(cand, root, pi_created(it), pi_created_open(it), 0u);
                                cand.node = vv;
                                cand.born = (x != 0) ? prog_stamp(true, rn, x) : 0u;
                                have = true;
                            }
                        } else if (active) {
                            const uint32_t kind = pi_kind(it), rc = pi_read_cell(it);
                            if (kind == kEdgeAny || (kind == kEdgeLit && pi_sym(it) == ch)) {  // :171-175
                                k3_working<NC>(cand, root, pi_created(it), pi_created_open(it),
                                               pi_prior_reads(it) & ~digit_bit);
                                cand.node = pi_node(it);
                                cand.born = prog_stamp(false, rn, x);
                                k3_apply<NC>(cand, pi_open(it), pi_close(it), i, 1u);
                                cand.first += 1;
                                have = true;
                            } else if (rc) {  // :176-193
                                const int k = int(rc) - 1;
                                const bool fresh = (pi_created(it) >> k) & 1u;
                                uint32_t L = 0, vs = 0;
#pragma unroll
                                for (int kk = 0; kk < NC; kk++)
                                    if (kk == k) {
                                        L = fresh ? 0u : root.len[kk];
                                        vs = root.start[kk];
                                    }
                                if (n32 - i >= L) {
                                    need_cmp = true;
                                    cmp_vs = vs;
                                    cmp_L = L;
                                    cand.born = prog_stamp(false, rn, x);
                                }
                            }
                        }
                    }
                }
                // backreference blocks: the whole tile compares each pending span; the loop runs
                // until the last tile of the warp has none left
                uint32_t cm = to.ballot(need_cmp);
                bool cmp_ok = false;
                while (TileOps<TILE>::warp_any(cm != 0u)) {
                    const int src = cm ? __ffs(int(cm)) - 1 : 0;
                    const uint32_t vs = __shfl_sync(ALL, cmp_vs, src, TILE);
                    uint32_t L = __shfl_sync(ALL, cmp_L, src, TILE);
                    if (!cm) L = 0;
                    bool eq;
                    if (!v.reversed) eq = tile_span_equal<TILE>(to, s + vs, s + i, L, lane);
                    else eq = tile_span_equal<TILE>(to, s + (n32 - vs - L), s + (n32 - i - L), L, lane);
                    // every lane waiting for this very span takes the answer (e.g. &1 read from two nodes)
                    const bool mine = need_cmp && cm && cmp_vs == vs && cmp_L == L;
                    if (mine) cmp_ok = eq;
                    cm &= ~to.ballot(mine);
                }
                if (need_cmp && cmp_ok) {
                    const uint32_t stamp = cand.born;
                    k3_working<NC>(cand, root, pi_created(it), pi_created_open(it),
                                   pi_prior_reads(it) & ~digit_bit);
                    cand.node = pi_node(it);
                    cand.born = stamp;
                    cand.first += cmp_L;
                    k3_apply<NC>(cand, pi_open(it), pi_close(it), i, cmp_L);
                    have = true;
                }
                // claim the target slot: smallest order key wins
                uint64_t k64 = 0;
                if (have) {
                    if (cand.first < i + 2u) next_near = true;  // its node will hold a configuration with first <= i+1
                    k64 = k3_key(cand.first, cand.flags, cand.born);
                    atomicMin(reinterpret_cast<unsigned long long *>(&KN[cand.node]), (unsigned long long)k64);
                }
                __syncwarp(ALL);
                if (have && KN[cand.node] == k64) {
                    const uint32_t slot = (cur ^ 1u) * SP + cand.node;
                    flg[slot] = cand.flags;
#pragma unroll
                    for (int k = 0; k < NC; k++) {
                        stt[slot * NC + k] = cand.start[k];
                        lnn[slot * NC + k] = cand.len[k];
                    }
                    // A configuration written in this pass may still be replaced in a later pass of the
                    // same step (T > TILE): the lanes then describe a superset of the new set, which can
                    // only withhold the jump or shorten it -- the step it lands on is run as usual.
                    const uint32_t pkey = (cand.node << gp.n_cells) | exists_mask_n<NC>(cand.flags);
                    if (cand.first < i + 2u || cand.first == n32 || s_begin[pkey] == 0xffffffffu ||
                        !(s_count[pkey] & kProgStable))
                        w_stable = false;
                    w_any = true;
                    if (cand.first < w_ev) w_ev = cand.first;
                    if (v.reversed) {
                        const uint32_t need = k3_need<NC>(cand.flags, cand.len);  // fresh: not yet tested against mfa.cpp:141
                        const uint32_t ps = need > n32 ? 0u : n32 - need + 1u;
                        if (ps < w_ev) w_ev = ps;
                    }
                }
                __syncwarp(ALL);
            }
            if (run) cur ^= 1u;  // states = new_states (:212)
            __syncwarp(ALL);
            next_near = to.any(next_near);
            bool jumped = false;
            if (run && (ovf || i == n32)) finished = true;
            // ---- C0. every configuration of the new set waits (first >= i + 2) and is reproduced
            //          unchanged by a step (kProgStable): the steps up to the first activation /
            //          reversed-mode pruning are the identity and are not run at all.  The facts come
            //          from the lanes that wrote the set in B; the slots are not read again ----
            const bool c0 = run && !finished && i + 2 < n32 && !next_near;
            if (TileOps<TILE>::warp_any(c0)) {
                const uint32_t nstable = to.ballot(!w_stable);  // three collectives, each by every tile
                const bool any = to.any(w_any);
                const uint32_t ev = to.min(w_ev);
                if (c0 && nstable == 0u && any && ev > i + 1) {
                    i = ev - 1;  // the increment below makes the next step ev
                    jumped = true;
                }
            }
            // ---- C. fast-forward over idle steps (see MfaSim::run); only after a step in which
            //         no configuration was active ----
            const bool c1 = run && !finished && !jumped && !any_active && i + 1 < n32;
            if (TileOps<TILE>::warp_any(c1)) {
                bool same = true;
                uint32_t ev = n32;
                for (uint32_t q = lane; q < SP; q += TILE) {
                    if (!c1) continue;
                    const uint64_t ka = keys[cur * SP + q], kb = keys[(cur ^ 1u) * SP + q];
                    const bool va = ka != K3_EMPTY, vb2 = kb != K3_EMPTY;
                    if (va != vb2) same = false;
                    else if (va) {
                        const uint32_t fa = uint32_t(ka >> 36), fb = uint32_t(kb >> 36);
                        const uint32_t sa = cur * SP + q, sb2 = (cur ^ 1u) * SP + q;
                        const uint32_t fla = flg[sa];
                        if (fa != fb || fla != flg[sb2] || fa <= i) same = false;
                        uint32_t need = 0;
#pragma unroll
                        for (int k = 0; k < NC; k++) {
                            const uint32_t fl = (fla >> (3 * k)) & 7u;
                            if (fl & 1u) {
                                const uint32_t la = lnn[sa * NC + k];
                                if (la != lnn[sb2 * NC + k] || (la && stt[sa * NC + k] != stt[sb2 * NC + k])) same = false;
                                if ((fl & 2u) || !(fl & 4u)) need += la;
                            }
                        }
                        if (fa < ev) ev = fa;
                        if (v.reversed && !(q == v.finish && fa == n32)) {
                            const uint32_t ps = n32 - need + 1u;
                            if (ps < ev) ev = ps;
                        }
                    }
                }
                same = to.all(same);
                ev = to.min(ev);
                // prog_stamp does not depend on the step index: further idle steps reproduce this
                // set bit for bit, so jump straight to the event step (the increment below adds 1)
                if (c1 && same && ev > i + 1) i = ev - 1;
            }
            if (run) i++;
        }
        if (finished) {
            if (lane == 0) {
                if (ovf) {
                    atomicAdd(overflow, 1ull);
                    out[si] = 0;
                } else {
                    out[si] = keys[cur * SP + v.finish] != K3_EMPTY ? 1 : 0;  // :230-235
                }
            }
            have_str = false;
        }
        __syncwarp(ALL);
    }
}

template <int NC, int TILE>
int launch_k3(const MfaView &v, const ProgView &gp, uint32_t n_items, uint32_t n_keys, const uint8_t *d_chars,
              Spans spans, const K1Rec *d_recs, uint64_t n, uint8_t *d_out, unsigned long long *d_overflow,
              unsigned long long *d_next, int sm_count, uint32_t sharing, cudaStream_t stream, const uint32_t *d_list,
              const unsigned long long *d_list_n, const uint32_t *d_gate, uint32_t gate_want) {
    constexpr int TILES = K3_WARPS * 32 / TILE;  // strings in flight per block
    const uint32_t SP = (v.n_states + TILE - 1u) & ~uint32_t(TILE - 1);
    const size_t per_tile = size_t(SP) * 2 * (8 + 4 + 8 * NC) + size_t(SP) * 12 + 16;
    const size_t tab = (size_t(n_keys) * 8 + 15) & ~size_t(15);
    const bool in_smem = size_t(n_items) * sizeof(ProgItem) <= 64 * 1024;
    const size_t smem = tab + (in_smem ? size_t(n_items) * sizeof(ProgItem) : 0) + TILES * per_tile;
    if (smem > 200 * 1024) return RXM_ERR_UNSUPPORTED;
    auto kern = k3_mfa_warp_kernel<NC, TILE>;
    // always the same value: handles that share a kernel instantiation may launch from several threads
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
        return RXM_ERR_CUDA;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, K3_WARPS * 32, smem) != cudaSuccess || nb <= 0)
        return RXM_ERR_CUDA;
    // rxm_set_concurrency: other handles' kernels run beside this one only if it leaves them block slots
    uint64_t blocks = uint64_t(sm_count) * nb / (sharing ? sharing : 1u);
    if (blocks == 0) blocks = 1;
    const uint64_t need = (n + TILES - 1) / TILES;
    if (blocks > need) blocks = need;
    if (cudaMemsetAsync(d_next, 0, sizeof(unsigned long long), stream) != cudaSuccess) return RXM_ERR_CUDA;
    RXM_LAUNCH(kern, unsigned(blocks), K3_WARPS * 32, smem, stream, v, gp, n_items, n_keys, in_smem ? 1u : 0u, d_chars, spans,
               d_recs, n, d_out, d_overflow, d_next, d_list, d_list_n, d_gate, gate_want);
    return RXM_OK;
}

template <int NC>
int launch_k3_tile(uint32_t tile, const MfaView &v, const ProgView &gp, uint32_t n_items, uint32_t n_keys,
                   const uint8_t *d_chars, Spans spans, const K1Rec *d_recs, uint64_t n, uint8_t *d_out,
                   unsigned long long *d_overflow, unsigned long long *d_next, int sm_count, uint32_t sharing, cudaStream_t stream,
                   const uint32_t *d_list, const unsigned long long *d_list_n, const uint32_t *d_gate, uint32_t gate_want) {
    // the per-string state (two buffers of one slot per node) of all strings of a block must fit
    // shared memory: automata with many nodes move to wider tiles (fewer strings per block)
    int st = RXM_ERR_UNSUPPORTED;
    if (tile <= 8) st = launch_k3<NC, 8>(v, gp, n_items, n_keys, d_chars, spans, d_recs, n, d_out, d_overflow, d_next, sm_count, sharing, stream, d_list, d_list_n, d_gate, gate_want);
    if (st == RXM_ERR_UNSUPPORTED && tile <= 16)
        st = launch_k3<NC, 16>(v, gp, n_items, n_keys, d_chars, spans, d_recs, n, d_out, d_overflow, d_next, sm_count, sharing, stream, d_list, d_list_n, d_gate, gate_want);
    if (st == RXM_ERR_UNSUPPORTED)
        st = launch_k3<NC, 32>(v, gp, n_items, n_keys, d_chars, spans, d_recs, n, d_out, d_overflow, d_next, sm_count, sharing, stream, d_list, d_list_n, d_gate, gate_want);
    return st;
}

}  // namespace

namespace {
// "long strings": mean length above 4096 -- the regime where one string per warp (K3, 32 lanes) beats one per
// thread (K4): few strings, each bounded by its own length.  One thread; flag[0] <- 1 / 0.
__global__ void mfa_pick_kernel(const Spans sp, uint64_t n, uint32_t *__restrict__ flag) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const uint64_t total = n ? sp.end[n - 1] - sp.begin[0] : 0;
        flag[0] = (n && total / n > kMfaLongMean) ? 1u : 0u;
    }
}
}  // namespace

int mfa_pick_launch(Spans spans, uint64_t n, uint32_t *d_flag, cudaStream_t stream) {
    RXM_LAUNCH(mfa_pick_kernel, 1u, 32, 0, stream, spans, n, d_flag);
    return RXM_OK;
}

int k3_launch(const MfaView &v, const ProgView &gp, uint32_t n_items, uint32_t n_keys, uint32_t n_cells,
              uint32_t tile, const uint8_t *d_chars, Spans spans, const K1Rec *d_recs, uint64_t n, uint8_t *d_out,
              unsigned long long *d_overflow, unsigned long long *d_next, int sm_count, uint32_t sharing,
              cudaStream_t stream, int *launched, const uint32_t *d_list, const unsigned long long *d_list_n,
              const uint32_t *d_gate, uint32_t gate_want) {
    *launched = 0;
    int st;
    if (n_cells <= 1) st = launch_k3_tile<1>(tile, v, gp, n_items, n_keys, d_chars, spans, d_recs, n, d_out, d_overflow, d_next, sm_count, sharing, stream, d_list, d_list_n, d_gate, gate_want);
    else if (n_cells <= 2) st = launch_k3_tile<2>(tile, v, gp, n_items, n_keys, d_chars, spans, d_recs, n, d_out, d_overflow, d_next, sm_count, sharing, stream, d_list, d_list_n, d_gate, gate_want);
    else if (n_cells <= 4) st = launch_k3_tile<4>(tile, v, gp, n_items, n_keys, d_chars, spans, d_recs, n, d_out, d_overflow, d_next, sm_count, sharing, stream, d_list, d_list_n, d_gate, gate_want);
    else return RXM_ERR_UNSUPPORTED;
    if (st == RXM_OK) *launched = 1;
    return st;
}

}  // namespace rxm
