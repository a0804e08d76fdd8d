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
                    std::vector<uint8_t> &accept, std::string *err, bool no_quad, bool no_oct) {
    kt = K1Tables();
    kt.n_states = p.n_states;
    kt.n_classes = p.n_classes;
    kt.start = p.start;
    kt.reversed = p.reversed;
    accept.assign(p.accept.begin(), p.accept.end());
    while (accept.size() % 256) accept.push_back(0);
    kt.accept_bytes = uint32_t(accept.size());
    int lit_lo = 256, lit_hi = -1;  // the window of the literal bytes
    for (int b = 0; b < 256; b++)
        if (p.byte_class[b]) {
            lit_lo = std::min(lit_lo, b);
            lit_hi = std::max(lit_hi, b);
        }
    // 65 - 128 sets over a two-letter window: the direct table has no room for a stride table (T and O / Q [128][256]
    // would be 64 KB of static shared memory next to 192 KB of rows), the two-lookup form has (Q[set][16] u16, 4 KB) --
    // measured on a 98-set automaton, 200 k strings: 0.194 ms per byte on the rows, 0.13 ms with four bytes per lookup
    const bool small_window_mid = p.n_states > 64 && lit_hi >= 0 && lit_hi - lit_lo <= 1 && !no_quad;
    if (p.n_states <= 128 && !small_window_mid) {  // direct table (static shared memory, <= 32 KB)
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
        if (lit_hi < 0) lit_lo = lit_hi = 'a';
        // Oct stride: with all literals in a TWO-letter window [lo, lo+1] a letter is one bit, and the
        // interior takes EIGHT input bytes per lookup: O[q][code], bit i of code = (byte i of the aligned
        // 8-byte group, memory order) - lo; composed in reading order like Q.
        if (p.n_states <= 64 && lit_hi - lit_lo <= 1 && !no_quad && !no_oct) {
            const uint32_t lo = uint32_t(std::min(lit_lo, 254));
            kt.quad = 2;
            kt.quad_lo = lo;
            const size_t base = table.size();
            table.resize(base + size_t(256) * sp, 0);
            for (uint32_t q = 0; q < p.n_states; q++)
                for (uint32_t code = 0; code < 256; code++) {
                    uint32_t r = q;
                    for (int k = 0; k < 8; k++) r = table[size_t(lo + ((code >> (p.reversed ? 7 - k : k)) & 1u)) * sp + r];
                    table[base + size_t(q) * 256 + code] = uint8_t(r);
                }
        } else if (p.n_states <= 64 && lit_hi - lit_lo <= 3 && !no_quad) {
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
        const size_t bytes = k1_classed_table_bytes(p.n_classes, p.n_states);
        if (!k1_classed_fits(p.n_classes, p.n_states)) {  // (plan_dfa stops there too)
            if (err) *err = "determinised automaton does not fit shared memory";
            return RXM_ERR_UNSUPPORTED;
        }
        table.assign((bytes + 15) & ~size_t(15), 0);
        for (uint32_t b = 0; b < 256; b++) table[b] = p.byte_class[b];
        uint16_t *tr = reinterpret_cast<uint16_t *>(table.data() + 256);
        for (size_t i = 0; i < p.trans.size(); i++) tr[i] = p.trans[i];
        // A stride for the two-lookup form: with all literals in a TWO-letter window [lo, lo+1] a letter is one bit,
        // and the interior of a string takes FOUR bytes per lookup from Q[set][16] u16 (32 bytes per set: up to ~4700
        // sets) if that still fits next to the per-byte tables, which stay for the first / last vectors and for vectors
        // with a byte outside the window.  Bit i of the code = (byte i of the aligned word, memory order) - lo; composed
        // in READING order.  (Eight bytes per lookup from O[set][256] u16 was built and measured: 512 bytes per set leave
        // one CTA per SM where Q leaves four -- 194 sets: 0.179 ms against 0.130 ms per 200 k strings.  Not kept.)
        if (lit_hi >= 0 && lit_hi - lit_lo <= 1 && !no_quad) {
            const uint32_t lo = uint32_t(std::min(lit_lo, 254));
            const size_t room = kK1ClassedBytes - kt.accept_bytes;
            const size_t base = table.size();
            if (base + size_t(32) * p.n_states <= room) {
                kt.quad = 1;
                kt.quad_lo = lo;
                kt.multi_off = uint32_t(base);
                table.resize(base + size_t(32) * p.n_states, 0);
                uint16_t *mt = reinterpret_cast<uint16_t *>(table.data() + base);
                tr = reinterpret_cast<uint16_t *>(table.data() + 256);  // (the vector moved)
                const uint32_t c0 = p.byte_class[lo], c1 = p.byte_class[lo + 1];
                for (uint32_t q = 0; q < p.n_states; q++)
                    for (uint32_t code = 0; code < 16; code++) {
                        uint32_t r = q;
                        for (int k = 0; k < 4; k++) {
                            const uint32_t bit = (code >> (p.reversed ? 3 - k : k)) & 1u;
                            r = tr[size_t(bit ? c1 : c0) * p.n_states + r];
                        }
                        mt[size_t(q) * 16 + code] = uint16_t(r);
                    }
            }
        }
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

// strings of 2^31-1 bytes and more are not scanned: their record carries this length, their bit is 0 (rxm.h)
constexpr uint32_t K1_LEN_OVERFLOW = 0x7fffffffu;
__device__ __forceinline__ uint32_t clamp_len(uint64_t len) { return len >= 0x7fffffffull ? K1_LEN_OVERFLOW : uint32_t(len); }

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
        uint32_t len[K1_SORT_PER_THREAD], bkt[K1_SORT_PER_THREAD], rank[K1_SORT_PER_THREAD];
        uint64_t fin[K1_SORT_PER_THREAD];
#pragma unroll
        for (int k = 0; k < K1_SORT_PER_THREAD; k++) {  // all loads first: one memory round trip, not one per string
            const uint64_t i = lo + uint32_t(k) * K1_SORT_THREADS + t;
            beg[k] = fin[k] = 0;
            if (i < n) {
                beg[k] = sp.begin[i];
                fin[k] = sp.end[i];
            }
        }
#pragma unroll
        for (int k = 0; k < K1_SORT_PER_THREAD; k++) {
            const uint64_t i = lo + uint32_t(k) * K1_SORT_THREADS + t;
            bkt[k] = 0xffffffffu;
            if (i < n) {
                const uint64_t l = fin[k] - beg[k];
                if (l >= 0x7fffffffull && overflow) atomicAdd(overflow, 1ull);
                len[k] = clamp_len(l);
                bkt[k] = len_bucket(len[k]);
                rank[k] = atomicAdd(&cnt[bkt[k]], 1u);  // its place among the strings of the bucket
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
            const uint32_t slot = cnt[bkt[k]] + rank[k];
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
// The same with an L2 eviction policy.  The row-staged scan reads every byte once: evict_first lets the streamed
// lines leave L2 ahead of the 256-byte blocks two neighbouring strings share (measured: -2.5 % per step).
__device__ __forceinline__ void cp_async16_pol(uint32_t smem_dst, const void *gsrc, bool pred, uint64_t pol) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t"
        "@p cp.async.cg.shared.global.L2::cache_hint.L2::256B [%0], [%1], 16, %3;\n\t}\n" ::"r"(smem_dst),
        "l"(gsrc), "r"(int(pred)), "l"(pol)
        : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
#ifdef RXM_TUNING  // tuning builds: the other L2 fetch granules, and an L2 prefetch of the line at p
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void cp_async16_128(uint32_t smem_dst, const void *gsrc, bool pred) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t"
        "@p cp.async.cg.shared.global.L2::128B [%0], [%1], 16;\n\t}\n" ::"r"(smem_dst),
        "l"(gsrc), "r"(int(pred))
        : "memory");
}
__device__ __forceinline__ void cp_async16_plain(uint32_t smem_dst, const void *gsrc, bool pred) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t"
        "@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}\n" ::"r"(smem_dst),
        "l"(gsrc), "r"(int(pred))
        : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p, bool pred) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %1, 0;\n\t"
        "@p prefetch.global.L2 [%0];\n\t}\n" ::"l"(p), "r"(int(pred))
        : "memory");
}
#endif
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
    if (pred && ((smem_dst | uint32_t(reinterpret_cast<uintptr_t>(gsrc))) & 15u)) simt::fail(4, "misaligned 16-byte cp.async");
    if (pred) memcpy(simt::shared_ptr(smem_dst), gsrc, 16);
}
inline void cp_async16_pol(uint32_t smem_dst, const void *gsrc, bool pred, uint64_t) { cp_async16(smem_dst, gsrc, pred); }
inline uint64_t policy_evict_first() { return 0; }
inline void cp_async_commit() {}
template <int N>
inline void cp_async_wait() {}
inline uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c) { return a * b + c; }
inline uint4 lds128(uint32_t addr) {
    if (addr & 15u) simt::fail(4, "misaligned 16-byte shared load");
    return *reinterpret_cast<const uint4 *>(simt::shared_ptr(addr));
}
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

// Oct stride (eight bytes per lookup; letters of a two-letter window are one bit each).  x = word - lo4
// leaves 0 / 1 in every byte (anything else shows under 0xFE in its own or a lower byte position);
// x0 * (2^3 + 2^10 + 2^17 + 2^24) puts the four bits of the first word at 24..27, x1 * (2^7 + 2^14 + 2^21 + 2^28)
// those of the second at 28..31, and no two partial products of the SUM share a bit position, so nothing
// carries: one IMAD with the other product as its addend.  Per 8 bytes: 2 IADD, 2 LOP, 2 IMAD, SHF off the
// dependent chain, IMAD + LDS.U8 on it.
struct OctStep {
    static constexpr bool on = true;
    const uint8_t *Q;   // shared-memory table O[q][code]
    uint32_t neg_lo4;
    __device__ __forceinline__ uint32_t pair(uint32_t q, uint32_t w0, uint32_t w1, uint32_t &bad) const {
        const uint32_t x0 = w0 + neg_lo4, x1 = w1 + neg_lo4;
        bad |= x0 | x1;
        const uint32_t code = mad_lo(x1, 0x10204080u, x0 * 0x01020408u) >> 24;
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
            q = pair(q, w[0], w[1], b);
            q = pair(q, w[2], w[3], b);
        } else {
            q = pair(q, w[2], w[3], b);
            q = pair(q, w[0], w[1], b);
        }
        bad = b & 0xfefefefeu;
        return q;
    }
};

// The quad stride for the two-lookup tables (u16 entries, one bit per letter of a two-letter window): four bytes per
// lookup from Q[set][16].
struct ClassedQuadStep {
    static constexpr bool on = true;
    const uint16_t *Q;
    uint32_t neg_lo4;
    __device__ __forceinline__ uint32_t word(uint32_t q, uint32_t w, uint32_t &bad) const {
        const uint32_t x = w + neg_lo4;
        bad |= x;
        return Q[mad_lo(q, 16u, (x * 0x01020408u) >> 24)];  // bits 24..27 <- the four letters, 28..31 stay clear
    }
    template <bool REV>
    __device__ __forceinline__ uint32_t vec(uint32_t q, const uint32_t (&w)[4], uint32_t &bad) const {
        uint32_t b = 0;
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
        bad = b & 0xfefefefeu;
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
            if (rec.len == K1_LEN_OVERFLOW) {  // reported by the tile sort; bit 0
                out[rec.idx] = 0;
                rec.len = 0;
                rec.idx = 0xffffffffu;
            }
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

// ---- row-staged scan (the product's scan for direct / quad / oct tables) ---------------------------------------
// What bounds the chunk-staged scan above is the shared-memory / L1 pipe, twice over (ncu): a
// cp.async instruction whose 32 lanes serve 8 strings touches 8-16 different 128-byte lines, and L1 works
// through one line per ~2 cycles; and 32 lanes looking up 32 unrelated table bytes meet in ~3.7 banks-deep
// conflicts.  The second is the stride's business (oct: half the lookups).  The first is this body's:
//   * a string is fetched in whole 128-byte LINES: 8 consecutive lanes copy the 8 16-byte pieces of one line,
//     one cp.async instruction serves 4 strings = 4 lines -- the least L1 can be asked for per 512 bytes --
//     and every sector crosses L2 -> SM once;
//   * a lane owns a ROW of STAGES lines in shared memory, used as a ring of 16-byte vectors.  Line k of its
//     string lands at ring vectors 8k .. 8k+7, rotated by a per-lane constant chosen so that the vector the
//     lane reads at step v sits in bank group (lane + v) mod 8: the eight lanes of an LDS.128 phase always
//     read eight different bank groups whatever the alignment of their strings (rows are a multiple of
//     128 bytes apart, no padding), and the 8 pieces of a line are written to 8 different groups;
//   * a string's first byte may stand anywhere in its first line, so a round of 8 vectors straddles lines
//     r and r+1: line r+1 must have landed before round r is walked, lines up to r+STAGES-1 are in flight.
// Reads: only 16-byte aligned pieces that hold at least one byte of the string.
template <bool REV, class Step, int STAGES, class Quad>
__device__ __forceinline__ void k1_rows_body(const Step st, const Quad qd, const uint8_t *__restrict__ chars,
                                             const K1Rec *__restrict__ recs, uint64_t n,
                                             uint8_t *__restrict__ out, const uint8_t *accept,
                                             uint32_t start_state, uint32_t *__restrict__ task_counter,
                                             uint32_t ring_base /* this warp's 32 rows, smem address */) {
    static_assert(STAGES >= 3 && STAGES <= 8, "a round reads two lines while at least one more is in flight");
    constexpr uint32_t RING = STAGES * 128u;  // bytes per row
    constexpr uint32_t R16 = RING / 16u;      // vectors per row
    constexpr int AHEAD = STAGES - 1;         // lines issued before the first round
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t sub = lane & 7u;                  // the 16-byte piece of a line this lane copies (ascending address)
    const uint32_t msub = REV ? 7u - sub : sub;      // ... counted in reading order
    const uint32_t row0 = ring_base + (lane >> 3) * RING;  // row of the string instruction 0 serves for this lane
    const uint32_t my_row = ring_base + lane * RING;
    const uint32_t ntiles = uint32_t((n + K1_TILE_STRINGS - 1) / K1_TILE_STRINGS);
    const uint64_t pol_first = policy_evict_first();
#if defined(RXM_TUNING) && defined(RXM_K1_HINT) && RXM_K1_HINT == 4
    const uint64_t pol_last = policy_evict_last();
#endif

    for (;;) {
        uint32_t task = 0;
        if (lane == 0) task = atomicAdd(task_counter, 1u);
        task = __shfl_sync(0xffffffffu, task, 0);
        // Tile-major: the 128 groups of a tile are handed out together, so the 256-byte blocks that neighbouring
        // strings share are fetched while their first reader's copy is still in L2 (-3.5 % per step against
        // "group g of every tile first"); within a tile the long groups come first.
#if defined(RXM_TUNING) && defined(RXM_K1_ORDER) && RXM_K1_ORDER == 0
        const uint32_t grp = task / ntiles, tile = task - grp * ntiles;
        if (grp >= K1_TILE_STRINGS / 32u) break;
#else
        const uint32_t tile = task / (K1_TILE_STRINGS / 32u), grp = task % (K1_TILE_STRINGS / 32u);
        if (tile >= ntiles) break;
#endif
        const uint64_t first = uint64_t(tile) * K1_TILE_STRINGS + grp * 32u;
        const uint64_t tile_end = min(n, uint64_t(tile + 1u) * K1_TILE_STRINGS);
        if (first >= tile_end) continue;

        K1Rec rec;
        rec.start = 0;
        rec.len = 0;
        rec.idx = 0xffffffffu;
        if (first + lane < tile_end) rec = recs[first + lane];
        if (rec.len == K1_LEN_OVERFLOW) {  // reported by the tile sort; bit 0
            out[rec.idx] = 0;
            rec.len = 0;
            rec.idx = 0xffffffffu;
        }
        uint32_t q = start_state;
        // The lane's stream: 16-byte vectors in reading order, vector 0 the aligned one that holds the first
        // byte read (hb pad bytes in front of it); `lead` vectors of line 0 come before vector 0.
        // org: forward the address of line 0, reversed the END of line 0 (lines then run downwards).
        uint64_t org;
        uint32_t lead, hb;
        {
            const uint64_t a = uint64_t(reinterpret_cast<uintptr_t>(chars + rec.start));
            if (!REV) {
                org = a & ~uint64_t(127);
                lead = uint32_t(a & 127u) >> 4;
                hb = uint32_t(a & 15u);
            } else {
                const uint64_t e = a + rec.len;
                org = (e + 127u) & ~uint64_t(127);
                const uint32_t d = uint32_t(org - e);
                lead = d >> 4;
                hb = d & 15u;
            }
        }
        const uint32_t nbytes = rec.len ? hb + rec.len : 0u;  // stream length incl. the front pad
        const uint32_t nvec = (nbytes + 15u) >> 4;
        const uint32_t rot = (lane - lead) & 7u;  // vector v of the stream lives at ring vector (rot + lead + v) mod R16
        // vectors [vlo, vhi) are complete (no pad, no tail) in EVERY stream of the warp
        const uint32_t vlo = __reduce_max_sync(0xffffffffu, hb ? 1u : 0u);
        const uint32_t vhi = __reduce_min_sync(0xffffffffu, nbytes >> 4);
        const uint32_t nvmax = __reduce_max_sync(0xffffffffu, nvec);

        // loader state: instruction g of a round serves string 4g + lane/8
        const uint8_t *src[8];  // this lane's piece of the next line to issue
        uint32_t hi[8];         // pieces (reading order, counted from line 0) below this index hold bytes of the string
        uint32_t pos[8];        // ring vector the piece goes to
        uint32_t lo_first[8];   // line 0 only: pieces below this index lie before the string
#pragma unroll
        for (int g = 0; g < 8; g++) {
            const int t = g * 4 + int(lane >> 3);
            const uint64_t o = __shfl_sync(0xffffffffu, org, t);
            const uint32_t ld = __shfl_sync(0xffffffffu, lead, t);
            const uint32_t nv = __shfl_sync(0xffffffffu, nvec, t);
            const uint32_t rt = __shfl_sync(0xffffffffu, rot, t);
            lo_first[g] = ld;
            hi[g] = nv ? ld + nv : 0u;
            pos[g] = rt + msub;  // < 15 <= R16
            src[g] = reinterpret_cast<const uint8_t *>(uintptr_t(!REV ? o + sub * 16u : o - 128u + sub * 16u));
        }
        uint32_t m = msub;  // reading-order index of this lane's piece in the next line to issue
        auto issue_line = [&](bool first_line) {
#pragma unroll
            for (int g = 0; g < 8; g++) {
                const bool want = m < hi[g] && (!first_line || m >= lo_first[g]);
#if defined(RXM_TUNING) && defined(RXM_K1_HINT)
#if RXM_K1_HINT == 0
                cp_async16_plain(row0 + uint32_t(g) * (4u * RING) + pos[g] * 16u, src[g], want);
#elif RXM_K1_HINT == 128
                cp_async16_128(row0 + uint32_t(g) * (4u * RING) + pos[g] * 16u, src[g], want);
#elif RXM_K1_HINT == 4  // the 256-byte blocks a string shares with its neighbours stay in L2, the others leave first
                const bool inner = m >= lo_first[g] + 16u && m + 16u < hi[g];
                cp_async16_pol(row0 + uint32_t(g) * (4u * RING) + pos[g] * 16u, src[g], want, inner ? pol_first : pol_last);
#elif RXM_K1_HINT == 6  // no eviction policy
                cp_async16(row0 + uint32_t(g) * (4u * RING) + pos[g] * 16u, src[g], want);
#else  // 256-byte granule only where the whole 256-byte block lies inside the string
                const bool inner = m >= lo_first[g] + 16u && m + 16u < hi[g];
                cp_async16(row0 + uint32_t(g) * (4u * RING) + pos[g] * 16u, src[g], want && inner);
                cp_async16_plain(row0 + uint32_t(g) * (4u * RING) + pos[g] * 16u, src[g], want && !inner);
#endif
#else
                cp_async16_pol(row0 + uint32_t(g) * (4u * RING) + pos[g] * 16u, src[g], want, pol_first);
#endif
                src[g] = !REV ? src[g] + 128 : src[g] - 128;
                pos[g] = pos[g] + 8u >= R16 ? pos[g] + 8u - R16 : pos[g] + 8u;
            }
            cp_async_commit();
            m += 8u;
        };

        issue_line(true);
#pragma unroll
        for (int s = 1; s < AHEAD; s++) issue_line(false);
        const uint32_t nrounds = (nvmax + 7u) >> 3;
        uint32_t pr = rot + lead;  // ring vector of the round's first vector
        for (uint32_t r = 0; r < nrounds; r++) {
            issue_line(false);  // line r + AHEAD (predicated off beyond each string's end; always commits)
#if defined(RXM_TUNING) && defined(RXM_K1_PF)
            {   // the lane's own string: line (r + AHEAD + PF) into L2
                const uint32_t kk = r + uint32_t(AHEAD) + uint32_t(RXM_K1_PF);
                const uint64_t pa = !REV ? org + uint64_t(kk) * 128u : org - uint64_t(kk + 1u) * 128u;
                prefetch_l2(reinterpret_cast<const void *>(uintptr_t(pa)), kk * 8u < lead + nvec && nvec != 0u);
            }
#endif
            cp_async_wait<AHEAD - 1>();  // lines <= r + 1 have landed
            __syncwarp();
            const uint32_t v0 = r * 8u;
            if (v0 >= vlo && v0 + 8u <= vhi) {
                // interior round: every vector complete in every stream
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    uint32_t pj = pr + uint32_t(j);
                    pj = pj >= R16 ? pj - R16 : pj;
                    const uint4 v = lds128(my_row + pj * 16u);
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                    if constexpr (Quad::on) {
                        uint32_t bad = 0;
                        const uint32_t qf = qd.template vec<REV>(q, w, bad);
                        if (bad) {  // a byte outside the window: this vector goes byte by byte
#pragma unroll
                            for (int k = 0; k < 16; k++) q = st(q, vec_byte<REV>(w, k));
                        } else {
                            q = qf;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 16; k++) q = st(q, vec_byte<REV>(w, k));
                    }
                }
            } else {
#pragma unroll 1
                for (int j = 0; j < 8; j++) {
                    const uint32_t lo = (v0 + uint32_t(j)) * 16u;  // stream offset of this vector
                    if (lo >= nbytes) continue;
                    uint32_t pj = pr + uint32_t(j);
                    pj = pj >= R16 ? pj - R16 : pj;
                    const uint4 v = lds128(my_row + pj * 16u);
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                    if (lo >= hb && lo + 16u <= nbytes) {
                        bool done = false;
                        if constexpr (Quad::on) {
                            uint32_t bad = 0;
                            const uint32_t qf = qd.template vec<REV>(q, w, bad);
                            if (!bad) {
                                q = qf;
                                done = true;
                            }
                        }
                        if (!done) {
#pragma unroll
                            for (int k = 0; k < 16; k++) q = st(q, vec_byte<REV>(w, k));
                        }
                    } else {
                        // boundary vector: byte k in READING order sits at stream offset lo + k
#pragma unroll
                        for (int k = 0; k < 16; k++) {
                            const uint32_t at = lo + uint32_t(k);
                            const uint32_t byte = vec_byte<REV>(w, k);
                            if (at >= hb && at < nbytes) q = st(q, byte);
                        }
                    }
                }
            }
            pr = pr + 8u >= R16 ? pr + 8u - R16 : pr + 8u;
            __syncwarp();
            // every live stream's active set is empty (automata.cpp:186-188)
            if (__all_sync(0xffffffffu, q == 0u || (r + 1u) * 128u >= nbytes)) break;
        }
        cp_async_wait<0>();
        __syncwarp();
        if (rec.idx != 0xffffffffu) out[rec.idx] = accept[q];
    }
}

// MODE 0: one byte per lookup (table T), 1: quad stride (T then Q), 2: oct stride (T then O)
template <bool REV, int L, int STAGES, int WARPS, int MODE>
__global__ void __launch_bounds__(WARPS * 32, 1)
k1_rows_kernel(const uint8_t *__restrict__ chars, const K1Rec *__restrict__ recs, uint64_t n,
               uint8_t *__restrict__ out, const uint8_t *__restrict__ g_table,
               const uint8_t *__restrict__ g_accept, uint32_t start, uint32_t quad_lo,
               uint32_t *__restrict__ task_counter) {
    constexpr uint32_t TB = 256u << L;
    constexpr uint32_t TABLES = MODE ? 2u * TB : TB;
    __shared__ __align__(16) uint8_t s_table[TABLES];  // static: offsets fold into the LDS
    __shared__ __align__(16) uint8_t s_accept[256];
    RXM_DYN_SMEM_128(ring);
    const uint4 *s4 = reinterpret_cast<const uint4 *>(g_table);
    uint4 *d4 = reinterpret_cast<uint4 *>(s_table);
    for (uint32_t i = threadIdx.x; i < TABLES / 16; i += blockDim.x) d4[i] = s4[i];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_accept[i] = g_accept[i];
    __syncthreads();
    const uint32_t ring0 = uint32_t(__cvta_generic_to_shared(ring)) + (threadIdx.x >> 5) * (32u * STAGES * 128u);
    const DirectStep<L> st{s_table};
    if constexpr (MODE == 2) {
        const OctStep qd{s_table + TB, 0u - quad_lo * 0x01010101u};
        k1_rows_body<REV, DirectStep<L>, STAGES, OctStep>(st, qd, chars, recs, n, out, s_accept, start, task_counter, ring0);
    } else if constexpr (MODE == 1) {
        const QuadStep qd{s_table + TB, 0u - quad_lo * 0x01010101u};
        k1_rows_body<REV, DirectStep<L>, STAGES, QuadStep>(st, qd, chars, recs, n, out, s_accept, start, task_counter, ring0);
    } else {
        k1_rows_body<REV, DirectStep<L>, STAGES, NoQuad>(st, NoQuad(), chars, recs, n, out, s_accept, start, task_counter, ring0);
    }
}

constexpr int K1_WARPS = 8;

#ifdef RXM_TUNING
// Chunk-staged ring with a multi-byte stride (round 1's scan): g_table = T[256][SP] followed by Q / O [SP][256].
template <bool REV, int L, int CH, int STAGES, int WARPS, class Multi>
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
    const Multi qd{s_table + TB, 0u - quad_lo * 0x01010101u};
    k1_scan_body<REV, DirectStep<L>, CH, STAGES, 1, Multi>(
        st, qd, chars, recs, n, out, s_accept, start, task_counter,
        ring0 + (threadIdx.x >> 5) * (STAGES * 32 * (CH + 16)));
}
#endif

// MODE 0: one byte per (two-step) lookup, 1: four bytes (Q at multi_off)
template <bool REV, int CH, int STAGES, int MODE>
__global__ void __launch_bounds__(K1_WARPS * 32)
k1_dfa_classed_kernel(const uint8_t *__restrict__ chars, const K1Rec *__restrict__ recs, uint64_t n,
                      uint8_t *__restrict__ out, const uint8_t *__restrict__ g_table, uint32_t table_bytes,
                      const uint8_t *__restrict__ g_accept, uint32_t accept_bytes, uint32_t n_states,
                      uint32_t start, uint32_t multi_off, uint32_t quad_lo, uint32_t *__restrict__ task_counter) {
    RXM_DYN_SMEM_128(smem);
    const uint4 *s4 = reinterpret_cast<const uint4 *>(g_table);
    uint4 *d4 = reinterpret_cast<uint4 *>(smem);
    for (uint32_t i = threadIdx.x; i < table_bytes / 16; i += blockDim.x) d4[i] = s4[i];
    for (uint32_t i = threadIdx.x; i < accept_bytes; i += blockDim.x) smem[table_bytes + i] = g_accept[i];
    __syncthreads();
    const uint32_t sbase = uint32_t(__cvta_generic_to_shared(smem));
    const uint32_t ring0 = (sbase + table_bytes + accept_bytes + 127u) & ~127u;
    const ClassedStep st{smem, reinterpret_cast<const uint16_t *>(smem + 256), n_states};
    const uint32_t my_ring = ring0 + (threadIdx.x >> 5) * (STAGES * 32 * (CH + 16));
    if constexpr (MODE == 1) {
        const ClassedQuadStep qd{reinterpret_cast<const uint16_t *>(smem + multi_off), 0u - quad_lo * 0x01010101u};
        k1_scan_body<REV, ClassedStep, CH, STAGES, 1, ClassedQuadStep>(st, qd, chars, recs, n, out, smem + table_bytes, start,
                                                                       task_counter, my_ring);
    } else {
        k1_scan_body<REV, ClassedStep, CH, STAGES, 1>(st, NoQuad(), chars, recs, n, out, smem + table_bytes, start, task_counter,
                                                      my_ring);
    }
}

// Geometry.  The product library holds ONE geometry per kernel: rows of 3 lines, 16 warps per CTA (one CTA per SM:
// 192 KB of rows) for the direct / quad / oct tables; the chunk-staged ring (64-byte chunks, 2 stages, 8 warps)
// for the two-lookup tables, whose size varies.  The others exist only in tuning builds (make EXTRA=-DRXM_TUNING),
// selected there with RXM_K1_VARIANT.
// The dynamic shared-memory limit is a per-FUNCTION attribute: it is always set to the same value, so that
// handles sharing a kernel instantiation may launch from several host threads (the size a launch needs is
// checked against it).
constexpr int K1_MAX_DYN_SMEM = 216 * 1024;
// ... the same value for every launch of a function: all the dynamic shared memory its static tables leave of the
// 227 KB a CTA may have (asking for more than that is an error at cudaFuncSetAttribute, which the automata with
// 33 - 128 sets -- 16 / 32 KB of static tables -- ran into: found by tests/fuzz/fuzz_tables_gpu.py)
template <class Kern>
int k1_set_dyn_smem(Kern kern, size_t need) {
#ifndef RXM_SIMT_HOST
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) return RXM_ERR_CUDA;
    const size_t room = size_t(227 * 1024) - fa.sharedSizeBytes;
    const size_t limit = room < size_t(K1_MAX_DYN_SMEM) ? room : size_t(K1_MAX_DYN_SMEM);
    if (need > limit) return RXM_ERR_UNSUPPORTED;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(limit)) != cudaSuccess)
        return RXM_ERR_CUDA;
#else
    (void)kern;
    (void)need;
#endif
    return RXM_OK;
}
struct V0 { static constexpr int CH = 64, STAGES = 2, NS = 1; };
#ifdef RXM_TUNING
inline int k1_variant() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("RXM_K1_VARIANT");
        v = e ? atoi(e) : 0;
        if (v < 0 || v > 15) v = 0;
    }
    return v;
}
#else
inline int k1_variant() { return 0; }
#endif

template <class Kern>
int blocks_per_sm(Kern kern, int threads, size_t smem) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem) != cudaSuccess) return 0;
    return nb;
}

template <bool REV, int L, int MODE, int STAGES, int WARPS>
int launch_rows_g(const K1Tables &kt, const K1Launch &a) {
    const size_t smem = size_t(WARPS) * 32 * STAGES * 128;
    static_assert(size_t(WARPS) * 32 * STAGES * 128 <= size_t(K1_MAX_DYN_SMEM), "rows fit the per-function limit");
    auto kern = k1_rows_kernel<REV, L, STAGES, WARPS, MODE>;
    if (const int st = k1_set_dyn_smem(kern, smem)) return st;
    const int nb = blocks_per_sm(kern, WARPS * 32, smem);
    if (nb <= 0) return RXM_ERR_CUDA;
    const uint64_t tasks = (a.n + 31) / 32;
    uint64_t blocks = uint64_t(a.sm_count) * nb;
    const uint64_t need = (tasks + WARPS - 1) / WARPS;
    if (blocks > need) blocks = need;
    RXM_LAUNCH(kern, unsigned(blocks), WARPS * 32, smem, a.stream, a.d_chars, a.d_recs, a.n, a.d_out, a.d_table, a.d_accept, kt.start, kt.quad_lo, a.d_task_counter);
    return RXM_OK;
}

#ifdef RXM_TUNING
template <bool REV, int L, int CH, int STAGES, int WARPS, class Multi>
int launch_chunks_w(const K1Tables &kt, const K1Launch &a) {
    const size_t smem = size_t(WARPS) * STAGES * 32 * (CH + 16);
    auto kern = k1_dfa_quad_kernel<REV, L, CH, STAGES, WARPS, Multi>;
    if (const int st = k1_set_dyn_smem(kern, smem)) return st;
    const int nb = blocks_per_sm(kern, WARPS * 32, smem);
    if (nb <= 0) return RXM_ERR_CUDA;
    const uint64_t tasks = (a.n + 31) / 32;
    uint64_t blocks = uint64_t(a.sm_count) * nb;
    const uint64_t need = (tasks + WARPS - 1) / WARPS;
    if (blocks > need) blocks = need;
    RXM_LAUNCH(kern, unsigned(blocks), WARPS * 32, smem, a.stream, a.d_chars, a.d_recs, a.n, a.d_out, a.d_table, a.d_accept, kt.start, kt.quad_lo, a.d_task_counter);
    return RXM_OK;
}
#endif

template <bool REV, int L, int MODE>
int launch_rows(const K1Tables &kt, const K1Launch &a) {
    switch (k1_variant()) {
#ifdef RXM_TUNING
        case 1: return launch_rows_g<REV, L, MODE, 4, 12>(kt, a);
        case 2: return launch_rows_g<REV, L, MODE, 4, 8>(kt, a);
        case 3: return launch_rows_g<REV, L, MODE, 3, 12>(kt, a);
        case 4: return launch_rows_g<REV, L, MODE, 3, 8>(kt, a);
        case 5: return launch_rows_g<REV, L, MODE, 5, 9>(kt, a);
        case 6: return launch_rows_g<REV, L, MODE, 6, 8>(kt, a);
        case 7: return launch_rows_g<REV, L, MODE, 3, 18>(kt, a);
        case 8:  // the chunk-staged ring of round 1 with this table's stride
            if constexpr (MODE == 2) return launch_chunks_w<REV, L, 64, 2, 8, OctStep>(kt, a);
            else if constexpr (MODE == 1) return launch_chunks_w<REV, L, 64, 2, 8, QuadStep>(kt, a);
            else return RXM_ERR_INVALID;
        case 9:
            if constexpr (MODE == 2) return launch_chunks_w<REV, L, 128, 2, 8, OctStep>(kt, a);
            else if constexpr (MODE == 1) return launch_chunks_w<REV, L, 128, 2, 8, QuadStep>(kt, a);
            else return RXM_ERR_INVALID;
#endif
        default: return launch_rows_g<REV, L, MODE, 3, 16>(kt, a);
    }
}

template <bool REV>
int launch_direct_l(const K1Tables &kt, const K1Launch &a) {
    if (kt.quad == 2) {
        switch (kt.log2sp) {
            case 4: return launch_rows<REV, 4, 2>(kt, a);
            case 5: return launch_rows<REV, 5, 2>(kt, a);
            case 6: return launch_rows<REV, 6, 2>(kt, a);
            default: return RXM_ERR_INVALID;
        }
    }
    if (kt.quad == 1) {
        switch (kt.log2sp) {
            case 4: return launch_rows<REV, 4, 1>(kt, a);
            case 5: return launch_rows<REV, 5, 1>(kt, a);
            case 6: return launch_rows<REV, 6, 1>(kt, a);
            default: return RXM_ERR_INVALID;
        }
    }
    switch (kt.log2sp) {
        case 4: return launch_rows<REV, 4, 0>(kt, a);
        case 5: return launch_rows<REV, 5, 0>(kt, a);
        case 6: return launch_rows<REV, 6, 0>(kt, a);
        case 7: return launch_rows<REV, 7, 0>(kt, a);
        default: return RXM_ERR_INVALID;
    }
}

template <bool REV, int MODE>
int launch_classed_m(const K1Tables &kt, const K1Launch &a) {
    constexpr size_t ring = size_t(K1_WARPS) * V0::STAGES * 32 * (V0::CH + 16);
    static_assert(kK1ClassedBytes + 128 + ring <= size_t(K1_MAX_DYN_SMEM), "the planner's table limit is what this launch can hold");
    const size_t smem = size_t(kt.table_bytes) + kt.accept_bytes + 128 + ring;
    auto kern = k1_dfa_classed_kernel<REV, V0::CH, V0::STAGES, MODE>;
    if (const int st = k1_set_dyn_smem(kern, smem)) return st;
    int nb = blocks_per_sm(kern, K1_WARPS * 32, smem);
    if (nb <= 0) return RXM_ERR_CUDA;
    const uint64_t tasks = (a.n + 31) / 32;
    uint64_t blocks = uint64_t(a.sm_count) * nb;
    const uint64_t need = (tasks + K1_WARPS - 1) / K1_WARPS;
    if (blocks > need) blocks = need;
    RXM_LAUNCH(kern, unsigned(blocks), K1_WARPS * 32, smem, a.stream, a.d_chars, a.d_recs, a.n, a.d_out, a.d_table, kt.table_bytes, a.d_accept, kt.accept_bytes, kt.n_states, kt.start, kt.multi_off, kt.quad_lo, a.d_task_counter);
    return RXM_OK;
}
template <bool REV>
int launch_classed(const K1Tables &kt, const K1Launch &a) {
    if (kt.quad == 1) return launch_classed_m<REV, 1>(kt, a);
    return launch_classed_m<REV, 0>(kt, a);
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
