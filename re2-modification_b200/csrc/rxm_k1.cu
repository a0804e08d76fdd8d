// K1 -- determinised memory-free automaton, sm_100a.
// Replaces Automata::match (automata.cpp:177-210) for whole batches.
//
// Data path (DESIGN.md "K1"):
//   1. tile sort     -- one kernel builds 16-byte records {start, len, index}; every tile of
//                       4096 strings is ordered by DESCENDING length bucket (counting sort
//                       in shared memory on a 1/64-granular log scale).  Strings are not moved.
//   2. scan kernel   -- persistent warps pull 32 consecutive records of a tile (== 32
//                       strings of nearly equal length) from an atomic task counter, the
//                       longest group of every tile first.  The warp streams the 32 strings in lock-step:
//                       quarter-/eighth-warps copy 16-byte pieces of each lane's next
//                       CH-byte chunk with cp.async.cg straight into a padded,
//                       bank-conflict-free shared-memory ring (no register staging,
//                       L1 bypassed, every global request a full aligned sector run);
//                       each lane then walks ITS string with one shared-memory table
//                       lookup per byte, state in a register.
//   Each lane's chunk grid is anchored at its string's 16-byte-aligned start, so only
//   the first and the last one or two 16-byte vectors of a string take the predicated
//   path; everything between runs the unpredicated 16-step body.
#include "rxm_kernels.cuh"

#include <algorithm>
#include <cstdlib>

namespace rxm {

int k1_build_tables(const DfaPlan &p, K1Tables &kt, std::vector<uint8_t> &table,
                    std::vector<uint8_t> &accept, std::string *err, bool no_quad) {
    kt = K1Tables();
    kt.n_states = p.n_states;
    kt.n_classes = p.n_classes;
    kt.start = p.start;
    kt.reversed = p.reversed;
    accept.assign(p.accept.begin(), p.accept.end());
    while (accept.size() % 256) accept.push_back(0);
    kt.accept_bytes = uint32_t(accept.size());
    if (p.n_states <= 128) {  // direct table (static shared memory, <= 32 KB)
        uint32_t l = 4;
        while ((1u << l) < p.n_states) l++;
        kt.mode = K1_DIRECT;
        kt.log2sp = l;
        const uint32_t sp = 1u << l;
        table.assign(size_t(256) * sp, 0);
        for (uint32_t b = 0; b < 256; b++)
            for (uint32_t s = 0; s < p.n_states; s++)
                table[size_t(b) * sp + s] = uint8_t(p.trans[size_t(p.byte_class[b]) * p.n_states + s]);
        // Quad stride: when every literal byte lies in one window [lo, lo+3] and the automaton
        // has <= 64 sets, the scan's interior takes FOUR input bytes per lookup:
        // Q[q][code], code = c3 | c2<<2 | c1<<4 | c0<<6 with ck = (byte k of the 32-bit word) - lo,
        // composed in READING order (right-to-left automata read byte 3 first).  A byte outside
        // the window sends its 16-byte vector to the per-byte table, so every input gets the
        // same answer as with the per-byte scan.
        int lit_lo = 256, lit_hi = -1;
        for (int b = 0; b < 256; b++)
            if (p.byte_class[b]) {
                lit_lo = std::min(lit_lo, b);
                lit_hi = std::max(lit_hi, b);
            }
        if (lit_hi < 0) lit_lo = lit_hi = 'a';
        if (p.n_states <= 64 && lit_hi - lit_lo <= 3 && !no_quad) {
            const uint32_t lo = uint32_t(std::min(lit_lo, 252));
            kt.quad = 1;
            kt.quad_lo = lo;
            const size_t base = table.size();
            table.resize(base + size_t(256) * sp, 0);
            for (uint32_t q = 0; q < p.n_states; q++)
                for (uint32_t code = 0; code < 256; code++) {
                    const uint32_t c[4] = {(code >> 6) & 3u, (code >> 4) & 3u, (code >> 2) & 3u, code & 3u};
                    uint32_t r = q;
                    for (int k = 0; k < 4; k++) r = table[size_t(lo + c[p.reversed ? 3 - k : k]) * sp + r];
                    table[base + size_t(q) * 256 + code] = uint8_t(r);
                }
        }
    } else {
        kt.mode = K1_CLASSED;
        const size_t bytes = 256 + 2 * size_t(p.n_classes) * p.n_states;
        if (bytes + kt.accept_bytes > 160 * 1024) {
            if (err) *err = "determinised automaton does not fit shared memory";
            return RXM_ERR_UNSUPPORTED;
        }
        table.assign((bytes + 15) & ~size_t(15), 0);
        for (uint32_t b = 0; b < 256; b++) table[b] = p.byte_class[b];
        uint16_t *tr = reinterpret_cast<uint16_t *>(table.data() + 256);
        for (size_t i = 0; i < p.trans.size(); i++) tr[i] = p.trans[i];
    }
    kt.table_bytes = uint32_t(table.size());
    return RXM_OK;
}

namespace {

// ---- bucket pass ---------------------------------------------------------------------
// Monotone length -> bucket map with <= 1/64 relative width: exact below 128, then six
// mantissa bits per octave (max index 128 + 24*64 - 1 < K1_BUCKETS).
__host__ __device__ __forceinline__ uint32_t len_bucket(uint32_t len) {
    if (len < 128u) return len;
#if defined(__CUDA_ARCH__)
    const uint32_t e = 31u - uint32_t(__clz(int(len)));
#else
    uint32_t e = 31;
    while (!((len >> e) & 1u)) e--;
#endif
    return 128u + (e - 7u) * 64u + ((len >> (e - 6u)) & 63u);
}

__device__ __forceinline__ uint32_t clamp_len(uint64_t len) { return len >= 0x7fffffffull ? 0u : uint32_t(len); }

// One kernel: every tile of K1_TILE consecutive strings is counting-sorted by DESCENDING length
// bucket inside shared memory and written out as K1_TILE records (strings are not moved).
// 32 consecutive records of a tile differ in length by ~1/64 of the tile's length range, which is
// all the scan needs for lock-step lanes; the scan's task order (group g of every tile before
// group g+1 of any) then runs the long groups of ALL tiles first without a global sort.
// Strings longer than 2^31-2 are reported and get bit 0.
constexpr int K1_TILE = int(K1_TILE_STRINGS);
constexpr int K1_SORT_THREADS = 1024;
constexpr int K1_SORT_PER_THREAD = K1_TILE / K1_SORT_THREADS;
__global__ void __launch_bounds__(K1_SORT_THREADS)
k1_tilesort_kernel(const Spans sp, uint64_t n, K1Rec *__restrict__ recs,
                   uint32_t *__restrict__ task_counter, unsigned long long *__restrict__ overflow /* may be null */) {
    static_assert(K1_BUCKETS == 2 * K1_SORT_THREADS, "two buckets per thread in the prefix pass");
    __shared__ uint32_t cnt[K1_BUCKETS];
    __shared__ uint32_t wsum[K1_SORT_THREADS / 32];
    const uint32_t t = threadIdx.x;
    if (blockIdx.x == 0 && t == 0) *task_counter = 0;
    const uint64_t ntiles = (n + K1_TILE - 1) / K1_TILE;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        cnt[t] = 0;
        cnt[t + K1_SORT_THREADS] = 0;
        __syncthreads();
        const uint64_t lo = tile * K1_TILE;
        uint64_t beg[K1_SORT_PER_THREAD];
        uint32_t len[K1_SORT_PER_THREAD], bkt[K1_SORT_PER_THREAD];
#pragma unroll
        for (int k = 0; k < K1_SORT_PER_THREAD; k++) {
            const uint64_t i = lo + uint32_t(k) * K1_SORT_THREADS + t;
            bkt[k] = 0xffffffffu;
            if (i < n) {
                beg[k] = sp.begin[i];
                const uint64_t l = sp.end[i] - beg[k];
                if (l >= 0x7fffffffull && overflow) atomicAdd(overflow, 1ull);
                len[k] = clamp_len(l);
                bkt[k] = len_bucket(len[k]);
                atomicAdd(&cnt[bkt[k]], 1u);
            }
        }
        __syncthreads();
        // exclusive prefix over the buckets, longest first: thread t owns buckets b0 > b1
        const uint32_t b0 = K1_BUCKETS - 1 - 2 * t, b1 = b0 - 1;
        const uint32_t c0 = cnt[b0], c1 = cnt[b1];
        uint32_t inc = c0 + c1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc, d);
            if ((t & 31u) >= uint32_t(d)) inc += v;
        }
        if ((t & 31u) == 31u) wsum[t >> 5] = inc;
        __syncthreads();
        if (t < 32) {
            uint32_t w = wsum[t];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, w, d);
                if (t >= uint32_t(d)) w += v;
            }
            wsum[t] = w;
        }
        __syncthreads();
        const uint32_t ex = inc - (c0 + c1) + ((t >> 5) ? wsum[(t >> 5) - 1] : 0u);
        cnt[b0] = ex;
        cnt[b1] = ex + c0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K1_SORT_PER_THREAD; k++) {
            if (bkt[k] == 0xffffffffu) continue;
            const uint32_t slot = atomicAdd(&cnt[bkt[k]], 1u);
            K1Rec r;
            r.start = beg[k];
            r.len = len[k];
            r.idx = uint32_t(lo + uint32_t(k) * K1_SORT_THREADS + t);
            recs[lo + slot] = r;
        }
        __syncthreads();
    }
}

// ---- scan kernel -----------------------------------------------------------------------
// 16-byte asynchronous copy, L1 bypassed; the L2::256B hint makes L2 fetch the whole 256-byte
// block on first touch, so DRAM sees 256-byte bursts per string instead of 64-byte ones
// (measured: 0.443 -> 0.418 ms per config-2 step; L2::128B changes nothing, .ca doubles the time).
#ifndef RXM_SIMT_HOST
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void *gsrc, bool pred) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t"
        "@p cp.async.cg.shared.global.L2::256B [%0], [%1], 16;\n\t}\n" ::"r"(smem_dst),
        "l"(gsrc), "r"(int(pred))
        : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c) {  // FMA pipe, not the ALU's
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
#else  // tests/hostsim: the copy lands at once (a legal outcome of the asynchronous one), shared addresses
       // are offsets into the emulated block's memory (simt_shim.hpp)
inline void cp_async16(uint32_t smem_dst, const void *gsrc, bool pred) {
    if (pred) memcpy(simt::shared_ptr(smem_dst), gsrc, 16);
}
inline void cp_async_commit() {}
template <int N>
inline void cp_async_wait() {}
inline uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c) { return a * b + c; }
inline uint4 lds128(uint32_t addr) { return *reinterpret_cast<const uint4 *>(simt::shared_ptr(addr)); }
#endif

// q' = T[byte][q] with T[byte][SP] u8 in shared memory.  The index is formed with one
// integer multiply-add (FMA pipe) so that the ALU pipe only carries the byte extraction.
template <int L>
struct DirectStep {
    const uint8_t *T;  // shared-memory table (address space resolved after inlining)
    __device__ __forceinline__ uint32_t operator()(uint32_t q, uint32_t byte) const {
        return T[mad_lo(byte, 1u << L, q)];
    }
};
struct ClassedStep {  // cmap[256] u8, then trans[class][n_states] u16
    const uint8_t *cmap;
    const uint16_t *trans;
    uint32_t n_states;
    __device__ __forceinline__ uint32_t operator()(uint32_t q, uint32_t byte) const {
        return trans[uint32_t(cmap[byte]) * n_states + q];
    }
};

template <bool REV, class Step>
__device__ __forceinline__ uint32_t step_word(const Step &st, uint32_t q, uint32_t w) {
    if (!REV) {
        q = st(q, __byte_perm(w, 0, 0x4440));
        q = st(q, __byte_perm(w, 0, 0x4441));
        q = st(q, __byte_perm(w, 0, 0x4442));
        q = st(q, __byte_perm(w, 0, 0x4443));
    } else {
        q = st(q, __byte_perm(w, 0, 0x4443));
        q = st(q, __byte_perm(w, 0, 0x4442));
        q = st(q, __byte_perm(w, 0, 0x4441));
        q = st(q, __byte_perm(w, 0, 0x4440));
    }
    return q;
}
template <bool REV, class Step>
__device__ __forceinline__ uint32_t step_vec(const Step &st, uint32_t q, const uint4 &v) {
    if (!REV) {
        q = step_word<REV>(st, q, v.x);
        q = step_word<REV>(st, q, v.y);
        q = step_word<REV>(st, q, v.z);
        q = step_word<REV>(st, q, v.w);
    } else {
        q = step_word<REV>(st, q, v.w);
        q = step_word<REV>(st, q, v.z);
        q = step_word<REV>(st, q, v.y);
        q = step_word<REV>(st, q, v.x);
    }
    return q;
}

// Quad stride (four bytes per lookup, DESIGN.md "K1").  x = word - lo4 maps the window's
// letters to 0..3 in every byte; any other byte leaves a bit under 0xFC in its own or a lower
// byte position and raises `bad`.  x * (1 + 2^10 + 2^20 + 2^30) gathers the four 2-bit codes
// into bits 24..31 (partial products land in distinct 2-bit slots, nothing carries), so the
// lookup address is q*256 + (prod >> 24): per word 3 instructions off the dependent chain and
// IMAD + LDS.U8 on it.
struct NoQuad {
    static constexpr bool on = false;
};
struct QuadStep {
    static constexpr bool on = true;
    const uint8_t *Q;   // shared-memory table Q[q][code]
    uint32_t neg_lo4;   // -(window base replicated into the four bytes)
    __device__ __forceinline__ uint32_t word(uint32_t q, uint32_t w, uint32_t &bad) const {
        const uint32_t x = w + neg_lo4;
        bad |= x;
        const uint32_t code = (x * 0x40100401u) >> 24;
        return Q[mad_lo(q, 256u, code)];
    }
    template <bool REV>
    __device__ __forceinline__ uint32_t vec(uint32_t q, const uint32_t (&w)[4], uint32_t &bad) const {
        uint32_t b = 0;
#ifdef RXM_K1_PROBE  // tuning builds only: the staging path without the lookups
        bad = 0;
        return q ^ ((w[0] ^ w[1] ^ w[2] ^ w[3]) & 1u);
#endif
        if (!REV) {
            q = word(q, w[0], b);
            q = word(q, w[1], b);
            q = word(q, w[2], b);
            q = word(q, w[3], b);
        } else {
            q = word(q, w[3], b);
            q = word(q, w[2], b);
            q = word(q, w[1], b);
            q = word(q, w[0], b);
        }
        bad = b & 0xfcfcfcfcu;
        return q;
    }
};

template <bool REV>
__device__ __forceinline__ uint32_t vec_byte(const uint32_t (&w)[4], int k) {  // k-th byte in READING order
    const int mb = REV ? 15 - k : k;
    return __byte_perm(w[mb >> 2], 0, 0x4440 + (mb & 3));
}

// A lane's chunk grid is anchored at the K1_ANCHOR-aligned block that holds the first byte it
// reads, so every CH-byte chunk is a whole number of 32-byte sectors (with 16-byte anchors half
// of the strings straddle three sectors per 64-byte chunk: 1.6x the bytes over the L2 -> SM
// crossbar, measured).  The pad in front (< K1_ANCHOR bytes) is skipped by the boundary path.
#ifndef RXM_K1_ANCHOR
#define RXM_K1_ANCHOR 32
#endif
constexpr uint32_t K1_ANCHOR = RXM_K1_ANCHOR;
static_assert(K1_ANCHOR == 16 || K1_ANCHOR == 32 || K1_ANCHOR == 64, "anchor alignment");

// CH = bytes per lane per stage, STAGES = ring depth, NS = strings walked by one lane at the
// same time (independent lookup chains: the LDS -> IMAD -> LDS chain of one string is ~34
// cycles per byte, so two interleaved chains double what a resident warp can issue).
// Lane stride CH+16 keeps the per-lane LDS.128 and the cooperative 16-byte cp.async writes
// bank-conflict free.
template <bool REV, class Step, int CH, int STAGES, int NS, class Quad = NoQuad>
__device__ __forceinline__ void k1_scan_body(const Step st, const Quad qd, const uint8_t *__restrict__ chars,
                                             const K1Rec *__restrict__ recs, uint64_t n,
                                             uint8_t *__restrict__ out, const uint8_t *accept,
                                             uint32_t start_state, uint32_t *__restrict__ task_counter,
                                             uint32_t ring_base /* this warp's ring, smem address */) {
    constexpr int LS = CH + 16;             // lane stride in the ring
    constexpr int STR_BYTES = 32 * LS;      // one string slot of every lane
    constexpr int STAGE_BYTES = NS * STR_BYTES;
    constexpr int VPC = CH / 16;            // vectors per chunk
    constexpr int LPT = CH / 16;            // lanes cooperating on one target lane's chunk
    constexpr int TPI = 32 / LPT;           // target lanes served per cp.async instruction
    constexpr int NI = 32 / TPI;            // cp.async instructions per round and string (== LPT)
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t part = lane % LPT;       // which 16-byte piece of the chunk this lane copies
    const uint32_t tsub = lane / LPT;       // which of the TPI targets
    const uint32_t my_ring = ring_base + lane * LS;
    const uint32_t ntiles = uint32_t((n + K1_TILE_STRINGS - 1) / K1_TILE_STRINGS);

    for (;;) {
        uint32_t task = 0;
        if (lane == 0) task = atomicAdd(task_counter, 1u);
        task = __shfl_sync(0xffffffffu, task, 0);
        // task -> (group of 32*NS records, tile): group g of every tile comes before group g+1
        const uint32_t grp = task / ntiles, tile = task - grp * ntiles;
        if (grp >= K1_TILE_STRINGS / (32u * NS)) break;
        const uint64_t first = uint64_t(tile) * K1_TILE_STRINGS + grp * (32u * NS);
        const uint64_t tile_end = min(n, uint64_t(tile + 1u) * K1_TILE_STRINGS);
        if (first >= tile_end) continue;

        // A lane's "stream" for string s: 16-byte vectors covering [p, p+len), anchored at the
        // aligned vector that holds the first byte read (forward: the string's first byte;
        // reversed: its last byte).  h = pad bytes in front of the stream.
        uint32_t idx[NS], h[NS], nbytes[NS], q[NS];
        const uint8_t *src[NS][NI];
        uint32_t src_lim[NS][NI];
        uint32_t dst_off[NI];
        uint32_t vlo = 0, vhi = 0xffffffffu, nvmax = 0;
#pragma unroll
        for (int s = 0; s < NS; s++) {
            K1Rec rec;
            rec.start = 0;
            rec.len = 0;
            rec.idx = 0xffffffffu;
            if (first + uint32_t(s) * 32u + lane < tile_end) rec = recs[first + uint32_t(s) * 32u + lane];
            idx[s] = rec.idx;
            q[s] = start_state;
            const uint8_t *p = chars + rec.start;
            const uint32_t len = rec.len;
            const uint8_t *anchor;  // forward: address of vector 0; reversed: END of vector 0
            if (!REV) {
                h[s] = uint32_t(reinterpret_cast<uintptr_t>(p)) & (K1_ANCHOR - 1u);
                anchor = p - h[s];
            } else {
                const uint8_t *e = p + len;
                h[s] = (K1_ANCHOR - (uint32_t(reinterpret_cast<uintptr_t>(e)) & (K1_ANCHOR - 1u))) & (K1_ANCHOR - 1u);
                anchor = e + h[s];
            }
            nbytes[s] = len ? h[s] + len : 0u;  // stream length incl. front pad
            const uint32_t nvec = (nbytes[s] + 15u) >> 4;
            // vectors [vlo, vhi) are complete (no pad, no tail) in EVERY stream of the warp
            vlo = max(vlo, __reduce_max_sync(0xffffffffu, (h[s] + 15u) >> 4));
            vhi = min(vhi, __reduce_min_sync(0xffffffffu, nbytes[s] >> 4));
            nvmax = max(nvmax, __reduce_max_sync(0xffffffffu, nvec));
            // sources / limits of the target lanes this lane copies for (round-invariant)
#pragma unroll
            for (int g = 0; g < NI; g++) {
                const uint32_t t = uint32_t(g) * TPI + tsub;
                const uint64_t a = __shfl_sync(0xffffffffu, uint64_t(reinterpret_cast<uintptr_t>(anchor)), int(t));
                const uint32_t nv = __shfl_sync(0xffffffffu, nvec, int(t));
                src_lim[s][g] = nv * 16u > part * 16u ? nv * 16u - part * 16u : 0u;  // copy while round offset < this
                if (!REV) src[s][g] = reinterpret_cast<const uint8_t *>(uintptr_t(a)) + part * 16u;
                else src[s][g] = reinterpret_cast<const uint8_t *>(uintptr_t(a)) - (part + 1u) * 16u;
                if (s == 0) dst_off[g] = ring_base + t * LS + part * 16u;
            }
        }
        const uint32_t nrounds = (nvmax + VPC - 1) / VPC;
        uint32_t issued = 0;  // rounds issued so far
        auto issue_round = [&]() {
            const uint32_t sb = (issued % STAGES) * STAGE_BYTES;
            const uint32_t boff = issued * CH;
#pragma unroll
            for (int s = 0; s < NS; s++) {
#pragma unroll
                for (int g = 0; g < NI; g++) {
                    cp_async16(dst_off[g] + sb + uint32_t(s) * STR_BYTES, src[s][g], boff < src_lim[s][g]);
                    src[s][g] = !REV ? src[s][g] + CH : src[s][g] - CH;
                }
            }
            cp_async_commit();
            issued++;
        };

#pragma unroll
        for (int s = 0; s < STAGES - 1; s++) issue_round();
        for (uint32_t r = 0; r < nrounds; r++) {
            issue_round();  // predicated off beyond each stream's end; always commits
            cp_async_wait<STAGES - 1>();
            __syncwarp();
            const uint32_t my = my_ring + (r % STAGES) * STAGE_BYTES;
            const uint32_t v0 = r * VPC;
            if (v0 >= vlo && v0 + VPC <= vhi) {
                // interior round: every vector complete in every stream; the NS chains interleave
#pragma unroll
                for (int j = 0; j < VPC; j++) {
                    uint32_t w[NS][4];
#pragma unroll
                    for (int s = 0; s < NS; s++) {
                        const uint4 v = lds128(my + uint32_t(s) * STR_BYTES + uint32_t(j) * 16u);
                        w[s][0] = v.x;
                        w[s][1] = v.y;
                        w[s][2] = v.z;
                        w[s][3] = v.w;
                    }
                    if constexpr (Quad::on) {
#pragma unroll
                        for (int s = 0; s < NS; s++) {
                            uint32_t bad = 0;
                            const uint32_t qf = qd.template vec<REV>(q[s], w[s], bad);
                            if (bad) {  // a byte outside the quad window: this vector goes byte by byte
#pragma unroll
                                for (int k = 0; k < 16; k++) q[s] = st(q[s], vec_byte<REV>(w[s], k));
                            } else {
                                q[s] = qf;
                            }
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 16; k++) {
#pragma unroll
                            for (int s = 0; s < NS; s++) q[s] = st(q[s], vec_byte<REV>(w[s], k));
                        }
                    }
                }
            } else {
#pragma unroll 1
                for (int j = 0; j < VPC; j++) {
                    const uint32_t lo = (v0 + uint32_t(j)) * 16u;  // stream offset of this vector
#pragma unroll
                    for (int s = 0; s < NS; s++) {
                        if (lo >= nbytes[s]) continue;
                        const uint4 v = lds128(my + uint32_t(s) * STR_BYTES + uint32_t(j) * 16u);
                        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                        if (lo >= h[s] && lo + 16u <= nbytes[s]) {
#pragma unroll
                            for (int k = 0; k < 16; k++) q[s] = st(q[s], vec_byte<REV>(w, k));
                        } else {
                            // boundary vector: byte k in READING order sits at stream offset lo + k
#pragma unroll
                            for (int k = 0; k < 16; k++) {
                                const uint32_t pos = lo + uint32_t(k);
                                const uint32_t byte = vec_byte<REV>(w, k);
                                if (pos >= h[s] && pos < nbytes[s]) q[s] = st(q[s], byte);
                            }
                        }
                    }
                }
            }
            __syncwarp();
            // every live stream's active set is empty (automata.cpp:186-188)
            bool idle = true;
#pragma unroll
            for (int s = 0; s < NS; s++) idle = idle && (q[s] == 0u || (r + 1u) * CH >= nbytes[s]);
            if (__all_sync(0xffffffffu, idle)) break;
        }
        cp_async_wait<0>();
        __syncwarp();
#pragma unroll
        for (int s = 0; s < NS; s++)
            if (idx[s] != 0xffffffffu) out[idx[s]] = accept[q[s]];
    }
}

constexpr int K1_WARPS = 8;

template <bool REV, int L, int CH, int STAGES, int NS>
__global__ void __launch_bounds__(K1_WARPS * 32)
k1_dfa_direct_kernel(const uint8_t *__restrict__ chars, const K1Rec *__restrict__ recs, uint64_t n,
                     uint8_t *__restrict__ out, const uint8_t *__restrict__ g_table,
                     const uint8_t *__restrict__ g_accept, uint32_t start,
                     uint32_t *__restrict__ task_counter) {
    constexpr uint32_t TB = 256u << L;
    __shared__ __align__(16) uint8_t s_table[TB];   // static: its offset folds into the LDS
    __shared__ __align__(16) uint8_t s_accept[256];
    RXM_DYN_SMEM_128(ring);
    const uint4 *s4 = reinterpret_cast<const uint4 *>(g_table);
    uint4 *d4 = reinterpret_cast<uint4 *>(s_table);
    for (uint32_t i = threadIdx.x; i < TB / 16; i += blockDim.x) d4[i] = s4[i];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_accept[i] = g_accept[i];
    __syncthreads();
    const uint32_t ring0 = uint32_t(__cvta_generic_to_shared(ring));
    const DirectStep<L> st{s_table};
    k1_scan_body<REV, DirectStep<L>, CH, STAGES, NS>(st, NoQuad(), chars, recs, n, out, s_accept, start, task_counter,
                                                     ring0 + (threadIdx.x >> 5) * (STAGES * NS * 32 * (CH + 16)));
}

// Quad-stride variant: g_table = T[256][SP] followed by Q[SP][256].
template <bool REV, int L, int CH, int STAGES, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k1_dfa_quad_kernel(const uint8_t *__restrict__ chars, const K1Rec *__restrict__ recs, uint64_t n,
                   uint8_t *__restrict__ out, const uint8_t *__restrict__ g_table,
                   const uint8_t *__restrict__ g_accept, uint32_t start, uint32_t quad_lo,
                   uint32_t *__restrict__ task_counter) {
    constexpr uint32_t TB = 256u << L;
    __shared__ __align__(16) uint8_t s_table[2 * TB];  // T then Q; static: offsets fold into the LDS
    __shared__ __align__(16) uint8_t s_accept[256];
    RXM_DYN_SMEM_128(ring);
    const uint4 *s4 = reinterpret_cast<const uint4 *>(g_table);
    uint4 *d4 = reinterpret_cast<uint4 *>(s_table);
    for (uint32_t i = threadIdx.x; i < 2 * TB / 16; i += blockDim.x) d4[i] = s4[i];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_accept[i] = g_accept[i];
    __syncthreads();
    const uint32_t ring0 = uint32_t(__cvta_generic_to_shared(ring));
    const DirectStep<L> st{s_table};
    const QuadStep qd{s_table + TB, 0u - quad_lo * 0x01010101u};
    k1_scan_body<REV, DirectStep<L>, CH, STAGES, 1, QuadStep>(
        st, qd, chars, recs, n, out, s_accept, start, task_counter,
        ring0 + (threadIdx.x >> 5) * (STAGES * 32 * (CH + 16)));
}

template <bool REV, int CH, int STAGES>
__global__ void __launch_bounds__(K1_WARPS * 32)
k1_dfa_classed_kernel(const uint8_t *__restrict__ chars, const K1Rec *__restrict__ recs, uint64_t n,
                      uint8_t *__restrict__ out, const uint8_t *__restrict__ g_table, uint32_t table_bytes,
                      const uint8_t *__restrict__ g_accept, uint32_t accept_bytes, uint32_t n_states,
                      uint32_t start, uint32_t *__restrict__ task_counter) {
    RXM_DYN_SMEM_128(smem);
    const uint4 *s4 = reinterpret_cast<const uint4 *>(g_table);
    uint4 *d4 = reinterpret_cast<uint4 *>(smem);
    for (uint32_t i = threadIdx.x; i < table_bytes / 16; i += blockDim.x) d4[i] = s4[i];
    for (uint32_t i = threadIdx.x; i < accept_bytes; i += blockDim.x) smem[table_bytes + i] = g_accept[i];
    __syncthreads();
    const uint32_t sbase = uint32_t(__cvta_generic_to_shared(smem));
    const uint32_t ring0 = (sbase + table_bytes + accept_bytes + 127u) & ~127u;
    const ClassedStep st{smem, reinterpret_cast<const uint16_t *>(smem + 256), n_states};
    k1_scan_body<REV, ClassedStep, CH, STAGES, 1>(st, NoQuad(), chars, recs, n, out, smem + table_bytes, start, task_counter,
                                                  ring0 + (threadIdx.x >> 5) * (STAGES * 32 * (CH + 16)));
}

// Ring geometry.  The product library holds ONE geometry per kernel (V0 / 64-byte chunks, 2 stages, 8 warps:
// the fastest measured); the others exist only in tuning builds (make EXTRA=-DRXM_TUNING),
// selected there with RXM_K1_VARIANT=0..7.
// The dynamic shared-memory limit is a per-FUNCTION attribute: it is always set to the same value, so that
// handles sharing a kernel instantiation may launch from several host threads (the size a launch needs is
// checked against it).
constexpr int K1_MAX_DYN_SMEM = 200 * 1024;
struct V0 { static constexpr int CH = 64, STAGES = 2, NS = 1; };
#ifdef RXM_TUNING
struct V1 { static constexpr int CH = 32, STAGES = 2, NS = 2; };
struct V2 { static constexpr int CH = 64, STAGES = 2, NS = 2; };
struct V3 { static constexpr int CH = 32, STAGES = 3, NS = 2; };
struct V4 { static constexpr int CH = 32, STAGES = 2, NS = 1; };

inline int k1_variant() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("RXM_K1_VARIANT");
        v = e ? atoi(e) : 0;
        if (v < 0 || v > 7) v = 0;
    }
    return v;
}
#else
inline int k1_variant() { return 0; }
#endif

template <class Kern>
int blocks_per_sm(Kern kern, size_t smem) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, K1_WARPS * 32, smem) != cudaSuccess) return 0;
    return nb;
}

template <bool REV, int L, class V>
int launch_direct_v(const K1Tables &kt, const K1Launch &a) {
    const size_t smem = size_t(K1_WARPS) * V::STAGES * V::NS * 32 * (V::CH + 16);
    auto kern = k1_dfa_direct_kernel<REV, L, V::CH, V::STAGES, V::NS>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_MAX_DYN_SMEM) != cudaSuccess)
        return RXM_ERR_CUDA;
    int nb = blocks_per_sm(kern, smem);
    if (nb <= 0) return RXM_ERR_CUDA;
    const uint64_t tasks = (a.n + 32 * V::NS - 1) / (32 * V::NS);
    uint64_t blocks = uint64_t(a.sm_count) * nb;
    const uint64_t need = (tasks + K1_WARPS - 1) / K1_WARPS;
    if (blocks > need) blocks = need;
    RXM_LAUNCH(kern, unsigned(blocks), K1_WARPS * 32, smem, a.stream, a.d_chars, a.d_recs, a.n, a.d_out, a.d_table, a.d_accept, kt.start, a.d_task_counter);
    return RXM_OK;
}

template <bool REV, int L>
int launch_direct(const K1Tables &kt, const K1Launch &a) {
    switch (k1_variant()) {
#ifdef RXM_TUNING
        case 1: return launch_direct_v<REV, L, V1>(kt, a);
        case 2: return launch_direct_v<REV, L, V2>(kt, a);
        case 3: return launch_direct_v<REV, L, V3>(kt, a);
        case 4: return launch_direct_v<REV, L, V4>(kt, a);
#endif
        default: return launch_direct_v<REV, L, V0>(kt, a);
    }
}

template <bool REV, int L, int CH, int STAGES, int WARPS>
int launch_quad_w(const K1Tables &kt, const K1Launch &a) {
    const size_t smem = size_t(WARPS) * STAGES * 32 * (CH + 16);
    auto kern = k1_dfa_quad_kernel<REV, L, CH, STAGES, WARPS>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_MAX_DYN_SMEM) != cudaSuccess)
        return RXM_ERR_CUDA;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, WARPS * 32, smem) != cudaSuccess || nb <= 0)
        return RXM_ERR_CUDA;
    const uint64_t tasks = (a.n + 31) / 32;
    uint64_t blocks = uint64_t(a.sm_count) * nb;
    const uint64_t need = (tasks + WARPS - 1) / WARPS;
    if (blocks > need) blocks = need;
    RXM_LAUNCH(kern, unsigned(blocks), WARPS * 32, smem, a.stream, a.d_chars, a.d_recs, a.n, a.d_out, a.d_table, a.d_accept, kt.start, kt.quad_lo, a.d_task_counter);
    return RXM_OK;
}

template <bool REV, int L>
int launch_quad(const K1Tables &kt, const K1Launch &a) {
    switch (k1_variant()) {  // tuning: CTA width / ring geometry
#ifdef RXM_TUNING
        case 1: return launch_quad_w<REV, L, 64, 2, 16>(kt, a);
        case 2: return launch_quad_w<REV, L, 64, 3, 10>(kt, a);
        case 3: return launch_quad_w<REV, L, 128, 2, 11>(kt, a);
        case 4: return launch_quad_w<REV, L, 128, 2, 8>(kt, a);
        case 5: return launch_quad_w<REV, L, 256, 2, 6>(kt, a);
        case 6: return launch_quad_w<REV, L, 128, 3, 7>(kt, a);
        case 7: return launch_quad_w<REV, L, 64, 4, 10>(kt, a);
#endif
        default: return launch_quad_w<REV, L, 64, 2, 8>(kt, a);
    }
}

template <bool REV>
int launch_direct_l(const K1Tables &kt, const K1Launch &a) {
    if (kt.quad) {
        switch (kt.log2sp) {
            case 4: return launch_quad<REV, 4>(kt, a);
            case 5: return launch_quad<REV, 5>(kt, a);
            case 6: return launch_quad<REV, 6>(kt, a);
            default: return RXM_ERR_INVALID;
        }
    }
    switch (kt.log2sp) {
        case 4: return launch_direct<REV, 4>(kt, a);
        case 5: return launch_direct<REV, 5>(kt, a);
        case 6: return launch_direct<REV, 6>(kt, a);
        case 7: return launch_direct<REV, 7>(kt, a);
        default: return RXM_ERR_INVALID;
    }
}

template <bool REV>
int launch_classed(const K1Tables &kt, const K1Launch &a) {
    const size_t smem = size_t(kt.table_bytes) + kt.accept_bytes + 128 + size_t(K1_WARPS) * V0::STAGES * 32 * (V0::CH + 16);
    auto kern = k1_dfa_classed_kernel<REV, V0::CH, V0::STAGES>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_MAX_DYN_SMEM) != cudaSuccess)
        return RXM_ERR_CUDA;
    int nb = blocks_per_sm(kern, smem);
    if (nb <= 0) return RXM_ERR_CUDA;
    const uint64_t tasks = (a.n + 31) / 32;
    uint64_t blocks = uint64_t(a.sm_count) * nb;
    const uint64_t need = (tasks + K1_WARPS - 1) / K1_WARPS;
    if (blocks > need) blocks = need;
    RXM_LAUNCH(kern, unsigned(blocks), K1_WARPS * 32, smem, a.stream, a.d_chars, a.d_recs, a.n, a.d_out, a.d_table, kt.table_bytes, a.d_accept, kt.accept_bytes, kt.n_states, kt.start, a.d_task_counter);
    return RXM_OK;
}

}  // namespace

int k1_tilesort_launch(Spans spans, uint64_t n, K1Rec *d_recs, uint32_t *d_counter, unsigned long long *d_overflow,
                       int sm_count, cudaStream_t stream) {
    const uint64_t ntiles = (n + K1_TILE - 1) / K1_TILE;
    uint64_t blocks = ntiles;
    const uint64_t cap = uint64_t(sm_count) * 2;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) return RXM_OK;
    RXM_LAUNCH(k1_tilesort_kernel, unsigned(blocks), K1_SORT_THREADS, 0, stream, spans, n, d_recs, d_counter, d_overflow);
    return RXM_OK;
}

int k1_launch(const K1Tables &kt, const K1Launch &a, int *launched) {
    *launched = 0;
    const uint64_t ntiles = (a.n + K1_TILE - 1) / K1_TILE;
    uint64_t blocks = ntiles;
    const uint64_t cap = uint64_t(a.sm_count) * 2;
    if (blocks > cap) blocks = cap;
    RXM_LAUNCH(k1_tilesort_kernel, unsigned(blocks), K1_SORT_THREADS, 0, a.stream, a.spans, a.n, a.d_recs, a.d_task_counter, a.d_overflow);
    *launched = 1;
    int st;
    if (kt.mode == K1_DIRECT) st = kt.reversed ? launch_direct_l<true>(kt, a) : launch_direct_l<false>(kt, a);
    else st = kt.reversed ? launch_classed<true>(kt, a) : launch_classed<false>(kt, a);
    if (st == RXM_OK) *launched = 2;
    return st;
}

}  // namespace rxm
