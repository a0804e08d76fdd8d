// Tokeniser -- whitespace-delimited text to string spans, sm_100a.
// The step BEFORE the hot path: the reference reads its inputs with
// `cin >> text` (matchers/match.cpp:22-23), i.e. tokens separated by runs of the C-locale
// whitespace bytes (space, \t, \n, \v, \f, \r); the token `exit` ends the input
// (match.cpp:24).  Doing that on the host costs more than matching on the device, so the raw
// text goes to HBM as it is and this pass finds the tokens there: ONE read of the text, 16
// bytes written per token.  Bound: HBM.
//
// Three kernels, no inter-block dependency (a one-pass chained scan was measured first: with
// ~1200 resident 16 KB pieces the decoupled look-back window never closes and 7 of 8 warps sit at
// the barrier -- 1.9 TB/s; ncu: barrier stall 22.8 per issue):
//   tok_classify  one coalesced read of the text; every 16-byte vector becomes a 16-bit
//                 whitespace mask (byte-parallel range tests, no per-byte loop; vectors without a
//                 candidate byte skip even those); masks (1 bit per input byte) and, per 2 KB warp
//                 piece, the (token starts, token ends) counts go to HBM
//   tok_offsets   exclusive scan of the per-block counts (one block; 8 bytes per 16 KB of text)
//   tok_emit      reads the MASKS only (1/8 of the text), recomputes start / end bits and writes
//                 begin[rank] / end[rank]; a warp whose piece holds no boundary leaves at once
// HBM traffic: text x 1.25 + 16 bytes per token.
#include "rxm_kernels.cuh"

namespace rxm {

namespace {

constexpr int TOK_THREADS = 256;
constexpr int TOK_BPT = 64;  // bytes per thread
constexpr uint64_t TOK_BLOCK_BYTES = uint64_t(TOK_THREADS) * TOK_BPT;

// bit 7 of every byte of the result: that byte of x is C-locale whitespace (9..13 or 32)
__device__ __forceinline__ uint32_t ws_bits(uint32_t x) {
    const uint32_t lo7 = x & 0x7f7f7f7fu;
    const uint32_t ge9 = lo7 + 0x77777777u;   // bit 7 <- (byte & 0x7f) >= 9
    const uint32_t gt13 = lo7 + 0x72727272u;  // bit 7 <- (byte & 0x7f) > 13
    const uint32_t ge32 = lo7 + 0x60606060u;  // bit 7 <- (byte & 0x7f) >= 32
    const uint32_t gt32 = lo7 + 0x5f5f5f5fu;  // bit 7 <- (byte & 0x7f) > 32
    return ((ge9 & ~gt13) | (ge32 & ~gt32)) & ~x & 0x80808080u;
}
// the four bit-7 flags of a word gathered into bits 0..3
__device__ __forceinline__ uint32_t gather4(uint32_t b7) { return (((b7 >> 7) * 0x00204081u) >> 21) & 0xfu; }

// (starts, ends) of one thread's 64 positions from its whitespace mask and the flag of the position before
__device__ __forceinline__ void tok_edges(uint64_t m, uint32_t prev, uint64_t &starts, uint64_t &ends) {
    const uint64_t pm = (m << 1) | prev;
    starts = ~m & pm;  // a token begins here
    ends = m & ~pm;    // a token ended just before here (exclusive end)
}

// 16-bit whitespace mask of a 16-byte vector.  Texts are mostly token bytes: one test for "any
// byte <= 0x20 or >= 0x80" (exact as an any-test, three operations per word) skips the per-byte
// classification for vectors without a candidate.
__device__ __forceinline__ uint32_t ws_mask16(const uint4 &v) {
    const uint32_t c = ((v.x - 0x21212121u) | v.x | (v.y - 0x21212121u) | v.y | (v.z - 0x21212121u) | v.z |
                        (v.w - 0x21212121u) | v.w) & 0x80808080u;
    if (c == 0) return 0;  // every byte in 0x21..0x7f: no borrow anywhere, no high bit
    return gather4(ws_bits(v.x)) | (gather4(ws_bits(v.y)) << 4) | (gather4(ws_bits(v.z)) << 8) |
           (gather4(ws_bits(v.w)) << 12);
}

// The unit of the scan is a WARP PIECE: 32 lanes x 64 positions = 2 KB of text.
//
// text0 = text rounded down to 16 bytes, pad = text - text0 (positions < pad are virtual
// whitespace); n0 = pad + nbytes.  Position n0 and everything after it is virtual whitespace
// too, so that a token that runs to the last byte still gets its end.  masks[] holds one bit per
// position (whitespace = 1) with those corrections applied; counts[b] = starts << 32 | ends of
// block b (16 KB), within[w] = the same for the pieces of w's block that precede piece w.
__global__ void __launch_bounds__(TOK_THREADS)
tok_classify_kernel(const uint8_t *__restrict__ text0, uint32_t pad, uint64_t n0, uint64_t *__restrict__ masks,
                    uint64_t *__restrict__ counts, uint64_t *__restrict__ within) {
    __shared__ __align__(8) uint16_t s_mask[TOK_THREADS * TOK_BPT / 16];  // one 16-bit mask per 16-byte vector
    __shared__ uint64_t s_tot[TOK_THREADS / 32];
    const uint32_t t = threadIdx.x, lane = t & 31u, bid = blockIdx.x;
    const uint64_t blk0 = uint64_t(bid) * TOK_BLOCK_BYTES, base = blk0 + uint64_t(t) * TOK_BPT;
    uint32_t prev_blk = 1u;  // whitespace flag of the byte before the block
    if (t == 0 && blk0 != 0 && blk0 - 1 < n0 && blk0 - 1 >= pad) {
        const uint32_t pb = text0[blk0 - 1];
        prev_blk = (pb == 32u || (pb - 9u) <= 4u) ? 1u : 0u;
    }
    {   // fully coalesced: thread t takes vectors t, t+256, ...
        const uint4 *src = reinterpret_cast<const uint4 *>(text0 + blk0);
        uint4 v[TOK_BPT / 16];
#pragma unroll
        for (int k = 0; k < TOK_BPT / 16; k++) {
            const uint64_t vo = blk0 + (uint64_t(k) * TOK_THREADS + t) * 16;
            v[k] = make_uint4(0x20202020u, 0x20202020u, 0x20202020u, 0x20202020u);  // past the end: whitespace
            if (vo < n0) v[k] = __ldg(src + k * TOK_THREADS + t);  // the last vector may run past n0 inside its 16-byte block
        }
#pragma unroll
        for (int k = 0; k < TOK_BPT / 16; k++) s_mask[k * TOK_THREADS + t] = uint16_t(ws_mask16(v[k]));
    }
    __syncthreads();
    uint64_t m = *reinterpret_cast<const uint64_t *>(&s_mask[4 * t]);  // this thread's 64 consecutive positions
    if (base < pad) m |= (pad - base >= 64) ? ~0ull : ((1ull << (pad - base)) - 1ull);
    if (base >= n0) m = ~0ull;
    else if (n0 - base < 64) m |= ~0ull << (n0 - base);
    masks[uint64_t(bid) * TOK_THREADS + t] = m;
    uint32_t prev = uint32_t(__shfl_up_sync(0xffffffffu, uint32_t(m >> 63), 1));
    if (lane == 0) {
        prev = (t == 0) ? prev_blk : uint32_t(s_mask[4 * t - 1] >> 15);
        if (t != 0 && (base - 1 < pad || base - 1 >= n0)) prev = 1u;  // s_mask lacks the corrections above
    }
    uint64_t starts, ends;
    tok_edges(m, prev, starts, ends);
    const uint32_t cs = __reduce_add_sync(0xffffffffu, uint32_t(__popcll(starts)));
    const uint32_t ce = __reduce_add_sync(0xffffffffu, uint32_t(__popcll(ends)));
    if (lane == 0) s_tot[t >> 5] = (uint64_t(cs) << 32) | ce;
    __syncthreads();
    if (lane == 0) {  // this piece's offset inside the block; the block total goes to the global scan
        uint64_t before = 0, all = 0;
#pragma unroll
        for (uint32_t w = 0; w < TOK_THREADS / 32; w++) {
            if (w < (t >> 5)) before += s_tot[w];
            all += s_tot[w];
        }
        within[uint64_t(bid) * (TOK_THREADS / 32) + (t >> 5)] = before;
        if (t == 0) counts[bid] = all;
    }
}

// counts[b] = (starts << 32 | ends) of block b  ->  exclusive prefix in place; result[0] = tokens in the
// text.  One block; a round takes 4096 entries (coalesced, 4 per thread) with two barriers.
constexpr int TOKO_THREADS = 1024, TOKO_SUB = 4;
__global__ void __launch_bounds__(TOKO_THREADS)
tok_offsets_kernel(uint64_t *__restrict__ counts, uint64_t n, unsigned long long *__restrict__ result) {
    __shared__ uint64_t s_w[TOKO_SUB * 32];  // inclusive totals of the 128 warp-tiles of a round, then their exclusive prefix
    __shared__ uint64_t s_carry;
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    if (t == 0) s_carry = 0;
    for (uint64_t b0 = 0; b0 < n; b0 += uint64_t(TOKO_THREADS) * TOKO_SUB) {
        uint64_t mine[TOKO_SUB], inc[TOKO_SUB];
#pragma unroll
        for (int k = 0; k < TOKO_SUB; k++) {
            const uint64_t i = b0 + uint64_t(k) * TOKO_THREADS + t;
            mine[k] = i < n ? counts[i] : 0ull;
            inc[k] = mine[k];
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
            for (int k = 0; k < TOKO_SUB; k++) {
                const uint64_t u = __shfl_up_sync(0xffffffffu, inc[k], d);
                if (int(lane) >= d) inc[k] += u;
            }
        }
        if (lane == 31) {
#pragma unroll
            for (int k = 0; k < TOKO_SUB; k++) s_w[k * 32 + warp] = inc[k];
        }
        __syncthreads();
        if (warp == 0) {  // 128 tile totals in entry order: tile (k, w) covers entries k*1024 + w*32 ..
            uint64_t x[TOKO_SUB], run = 0;
            // read before the shuffles below, which order it against lane 31's rewrite after them (found by
            // the SIMT emulator's shuffled lane order: nothing else keeps the lanes of a warp in step)
            const uint64_t carry_in = s_carry;
#pragma unroll
            for (int k = 0; k < TOKO_SUB; k++) x[k] = s_w[lane * TOKO_SUB + k];  // lane owns tiles 4*lane .. 4*lane+3
#pragma unroll
            for (int k = 0; k < TOKO_SUB; k++) run += x[k];
            uint64_t incl = run;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t u = __shfl_up_sync(0xffffffffu, incl, d);
                if (int(lane) >= d) incl += u;
            }
            uint64_t ex = carry_in + incl - run;
#pragma unroll
            for (int k = 0; k < TOKO_SUB; k++) {
                s_w[lane * TOKO_SUB + k] = ex;
                ex += x[k];
            }
            if (lane == 31) s_carry = ex;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TOKO_SUB; k++) {
            const uint64_t i = b0 + uint64_t(k) * TOKO_THREADS + t;
            if (i < n) counts[i] = s_w[k * 32 + warp] + inc[k] - mine[k];
        }
        __syncthreads();  // s_w is rewritten by the next round
    }
    if (t == 0) result[0] = s_carry >> 32;
}

// One warp per TOKE_PIECES consecutive warp pieces (their mask loads are issued together), no
// block-level step: most pieces hold no token boundary at all and are done after a ballot.
constexpr int TOKE_PIECES = 4;
__global__ void __launch_bounds__(TOK_THREADS)
tok_emit_kernel(const uint64_t *__restrict__ masks, const uint64_t *__restrict__ bases,
                const uint64_t *__restrict__ within, uint64_t nwarps, uint32_t pad,
                uint64_t *__restrict__ begin, uint64_t *__restrict__ end, uint64_t cap) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t w0 = (uint64_t(blockIdx.x) * (TOK_THREADS / 32) + (threadIdx.x >> 5)) * TOKE_PIECES;
    if (w0 >= nwarps) return;
    uint64_t m[TOKE_PIECES];
#pragma unroll
    for (int j = 0; j < TOKE_PIECES; j++) m[j] = (w0 + j < nwarps) ? masks[(w0 + j) * 32 + lane] : ~0ull;
    uint32_t carry = 1u;  // whitespace flag of the position before piece j
    if (lane == 0 && w0 != 0) carry = uint32_t(masks[w0 * 32 - 1] >> 63);
    carry = __shfl_sync(0xffffffffu, carry, 0);
#pragma unroll
    for (int j = 0; j < TOKE_PIECES; j++) {
        if (!carry && !__any_sync(0xffffffffu, m[j] != 0)) continue;  // 2 KB inside one token: no boundary
        const uint32_t top = uint32_t(m[j] >> 63);
        uint32_t prev = __shfl_up_sync(0xffffffffu, top, 1);
        if (lane == 0) prev = carry;
        carry = __shfl_sync(0xffffffffu, top, 31);
        uint64_t starts, ends;
        tok_edges(m[j], prev, starts, ends);
        if (w0 + j >= nwarps || !__any_sync(0xffffffffu, (starts | ends) != 0)) continue;
        const uint32_t mine = (uint32_t(__popcll(starts)) << 16) | uint32_t(__popcll(ends));  // <= 32 each per lane
        uint32_t inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d);
            if (int(lane) >= d) inc += u;
        }
        const uint32_t ex = inc - mine;
        const uint64_t w = w0 + j;
        const uint64_t bb = bases[w / (TOK_THREADS / 32)] + within[w];
        uint64_t rs = (bb >> 32) + (ex >> 16), re = (bb & 0xffffffffull) + (ex & 0xffffu);
        const uint64_t base = (w * 32 + lane) * TOK_BPT;
        uint64_t b = starts;
        while (b) {
            const int k = __ffsll(static_cast<long long>(b)) - 1;
            b &= b - 1;
            if (rs < cap) begin[rs] = base + uint64_t(k) - pad;
            rs++;
        }
        b = ends;
        while (b) {
            const int k = __ffsll(static_cast<long long>(b)) - 1;
            b &= b - 1;
            if (re < cap) end[re] = base + uint64_t(k) - pad;
            re++;
        }
    }
}

// result[1] <- index of the first token that is exactly "exit" (match.cpp:24)
__global__ void __launch_bounds__(256)
tok_exit_kernel(const uint8_t *__restrict__ text, const uint64_t *__restrict__ begin, const uint64_t *__restrict__ end,
                uint64_t cap, unsigned long long *__restrict__ result) {
    const uint64_t n = min(uint64_t(result[0]), cap);
    for (uint64_t k = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += uint64_t(gridDim.x) * blockDim.x) {
        const uint64_t b = begin[k];
        if (end[k] - b != 4) continue;
        if (text[b] == 'e' && text[b + 1] == 'x' && text[b + 2] == 'i' && text[b + 3] == 't')
            atomicMin(result + 1, static_cast<unsigned long long>(k));
    }
}

}  // namespace

uint64_t tok_blocks(uint64_t nbytes) { return (nbytes + 16 + TOK_BLOCK_BYTES) / TOK_BLOCK_BYTES + 1; }

int tok_launch(const uint8_t *d_text, uint64_t nbytes, uint64_t *d_begin, uint64_t *d_end, uint64_t cap,
               const TokWork &w, int sm_count, cudaStream_t stream, int *launched) {
    *launched = 0;
    if (nbytes >= (1ull << 32) - 64) return RXM_ERR_UNSUPPORTED;  // per-call counts are 32 bits
    const uint32_t pad = uint32_t(reinterpret_cast<uintptr_t>(d_text)) & 15u;
    const uint64_t n0 = nbytes + pad;
    const uint64_t nblocks = (n0 + 1 + TOK_BLOCK_BYTES - 1) / TOK_BLOCK_BYTES;  // position n0 included
    if (nblocks > w.blocks_cap) return RXM_ERR_INVALID;
    if (cudaMemsetAsync(w.d_result + 1, 0xff, sizeof(unsigned long long), stream) != cudaSuccess) return RXM_ERR_CUDA;
    const uint64_t nwarps = nblocks * (TOK_THREADS / 32);
    uint64_t *d_within = w.d_counts + w.blocks_cap;  // [blocks_cap * 8] after the block totals
    RXM_LAUNCH(tok_classify_kernel, unsigned(nblocks), TOK_THREADS, 0, stream, d_text - pad, pad, n0, w.d_masks, w.d_counts, d_within);
    RXM_LAUNCH(tok_offsets_kernel, 1, TOKO_THREADS, 0, stream, w.d_counts, nblocks, w.d_result);
    const uint64_t eblocks = (nwarps + TOKE_PIECES * (TOK_THREADS / 32) - 1) / (TOKE_PIECES * (TOK_THREADS / 32));
    RXM_LAUNCH(tok_emit_kernel, unsigned(eblocks), TOK_THREADS, 0, stream, w.d_masks, w.d_counts, d_within, nwarps, pad, d_begin, d_end, cap);
    *launched = 3;
    uint64_t eb = (cap + 255) / 256;
    const uint64_t ecap = uint64_t(sm_count) * 8;
    if (eb > ecap) eb = ecap;
    if (eb == 0) eb = 1;
    RXM_LAUNCH(tok_exit_kernel, unsigned(eb), 256, 0, stream, d_text, d_begin, d_end, cap, w.d_result);
    *launched = 4;
    return RXM_OK;
}

}  // namespace rxm
