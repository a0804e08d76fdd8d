// Tokeniser -- whitespace-delimited text to string spans, sm_100a.
// The step BEFORE the hot path: the reference reads its inputs with
// `cin >> text` (matchers/match.cpp:22-23), i.e. tokens separated by runs of the C-locale
// whitespace bytes (space, \t, \n, \v, \f, \r); the token `exit` ends the input
// (match.cpp:24).  Doing that on the host costs more than matching on the device, so the raw
// text goes to HBM as it is and this pass finds the tokens there: ONE read of the text, 16
// bytes written per token.  Bound: HBM.
//
// One pass, chained scan: a block takes the next 16 KB piece (atomic ticket, so earlier
// pieces are always running), every thread classifies its 64 bytes into a 64-bit whitespace
// mask (byte-parallel range tests, no per-byte loop), derives the token-start and token-end
// bits from the mask and the previous position, the block publishes its (starts, ends) counts
// in one 64-bit status word and looks back for its prefix (decoupled look-back, warp-wide),
// and the threads write begin[rank] / end[rank] straight from their mask registers.
#include "rxm_kernels.cuh"

namespace rxm {

namespace {

constexpr int TOK_THREADS = 256;
constexpr int TOK_BPT = 64;  // bytes per thread
constexpr uint64_t TOK_BLOCK_BYTES = uint64_t(TOK_THREADS) * TOK_BPT;

// status word: flag:2 | starts:31 | ends:31
constexpr uint64_t TOK_AGG = 1ull << 62, TOK_INC = 2ull << 62, TOK_FLAGS = 3ull << 62;
__device__ __forceinline__ uint64_t tok_pack(uint64_t flag, uint64_t s, uint64_t e) { return flag | (s << 31) | e; }

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

__device__ __forceinline__ uint64_t ld_status(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_status(uint64_t *p, uint64_t v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// text0 = text rounded down to 16 bytes, pad = text - text0 (positions < pad are virtual
// whitespace); n0 = pad + nbytes.  Position n0 is virtual whitespace too, so that a token that
// runs to the last byte still gets its end.
__global__ void __launch_bounds__(TOK_THREADS)
tok_scan_kernel(const uint8_t *__restrict__ text0, uint32_t pad, uint64_t n0, uint64_t *__restrict__ begin,
                uint64_t *__restrict__ end, uint64_t cap, uint64_t *__restrict__ status,
                uint32_t *__restrict__ ticket, unsigned long long *__restrict__ result, uint32_t nblocks) {
    __shared__ uint32_t s_bid;
    __shared__ uint32_t s_warp[TOK_THREADS / 32];
    __shared__ uint64_t s_prefix;
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    if (t == 0) s_bid = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t bid = s_bid;
    if (bid >= nblocks) return;
    const uint64_t base = uint64_t(bid) * TOK_BLOCK_BYTES + uint64_t(t) * TOK_BPT;

    // ---- classify: 64-bit whitespace mask of positions base .. base+63 ----
    uint64_t m = ~0ull;
    if (base < n0) {
        m = 0;
        const uint4 *src = reinterpret_cast<const uint4 *>(text0 + base);
#pragma unroll
        for (int k = 0; k < TOK_BPT / 16; k++) {
            uint32_t nib = 0xffffu;
            if (base + uint64_t(k) * 16 < n0) {  // the last vector may run past n0 inside its 16-byte block
                const uint4 v = __ldg(src + k);
                nib = gather4(ws_bits(v.x)) | (gather4(ws_bits(v.y)) << 4) | (gather4(ws_bits(v.z)) << 8) |
                      (gather4(ws_bits(v.w)) << 12);
            }
            m |= uint64_t(nib) << (16 * k);
        }
        if (base < pad) m |= (pad - base >= 64) ? ~0ull : ((1ull << (pad - base)) - 1ull);
        if (n0 - base < 64) m |= ~0ull << (n0 - base);
    }
    // whitespace flag of position base-1: the previous thread's top bit
    uint32_t prev = uint32_t(__shfl_up_sync(0xffffffffu, uint32_t(m >> 63), 1));
    if (lane == 0) {
        prev = 1u;
        if (base != 0 && base - 1 < n0 && base - 1 >= pad) {
            const uint32_t b = text0[base - 1];
            prev = (b == 32u || (b - 9u) <= 4u) ? 1u : 0u;
        }
    }
    const uint64_t pm = (m << 1) | prev;
    const uint64_t starts = ~m & pm;  // token begins here
    const uint64_t ends = m & ~pm;    // token ended just before here (exclusive end)
    const uint32_t cs = uint32_t(__popcll(starts)), ce = uint32_t(__popcll(ends));

    // ---- block scan of (cs, ce), packed 16:16 (a block holds < 2^14 of either) ----
    uint32_t inc = (cs << 16) | ce;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d);
        if (int(lane) >= d) inc += u;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t woff = 0, btot = 0;
#pragma unroll
    for (int w = 0; w < TOK_THREADS / 32; w++) {
        const uint32_t x = s_warp[w];
        if (w < int(warp)) woff += x;
        btot += x;
    }
    const uint32_t ex = inc - ((cs << 16) | ce) + woff;  // exclusive, inside the block

    // ---- chained scan over blocks: publish the aggregate, look back for the prefix ----
    if (warp == 0) {
        const uint64_t bs = btot >> 16, be = btot & 0xffffu;
        if (lane == 0) st_status(status + bid, tok_pack(bid == 0 ? TOK_INC : TOK_AGG, bs, be));
        uint64_t ps = 0, pe = 0;
        if (bid != 0) {
            int64_t j = int64_t(bid) - 1;
            for (;;) {
                const int64_t mine = j - int64_t(lane);
                uint64_t w = TOK_INC;  // blocks before the first: an empty inclusive prefix
                if (mine >= 0) {
                    do w = ld_status(status + mine);
                    while ((w & TOK_FLAGS) == 0);
                }
                const uint32_t incl = __ballot_sync(0xffffffffu, (w & TOK_FLAGS) == TOK_INC);
                const int stop = incl ? __ffs(int(incl)) - 1 : 32;  // nearest block with an inclusive prefix
                uint64_t s = (int(lane) <= stop) ? ((w >> 31) & 0x7fffffffull) : 0ull;
                uint64_t e = (int(lane) <= stop) ? (w & 0x7fffffffull) : 0ull;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    s += __shfl_xor_sync(0xffffffffu, s, d);
                    e += __shfl_xor_sync(0xffffffffu, e, d);
                }
                ps += s;
                pe += e;
                if (incl) break;
                j -= 32;
            }
            if (lane == 0) st_status(status + bid, tok_pack(TOK_INC, ps + bs, pe + be));
        }
        if (lane == 0) {
            s_prefix = (ps << 32) | pe;
            if (bid == nblocks - 1) result[0] = ps + bs;  // tokens in the whole text
        }
    }
    __syncthreads();
    uint64_t rs = (s_prefix >> 32) + (ex >> 16), re = (s_prefix & 0xffffffffull) + (ex & 0xffffu);

    // ---- emit ----
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

uint64_t tok_status_words(uint64_t nbytes) { return (nbytes + 16 + TOK_BLOCK_BYTES) / TOK_BLOCK_BYTES + 1; }

int tok_launch(const uint8_t *d_text, uint64_t nbytes, uint64_t *d_begin, uint64_t *d_end, uint64_t cap,
               const TokWork &w, int sm_count, cudaStream_t stream, int *launched) {
    *launched = 0;
    if (nbytes >= (1ull << 32) - 64) return RXM_ERR_UNSUPPORTED;  // counts are 31 bits per status word
    const uint32_t pad = uint32_t(reinterpret_cast<uintptr_t>(d_text)) & 15u;
    const uint64_t n0 = nbytes + pad;
    const uint64_t nblocks = (n0 + 1 + TOK_BLOCK_BYTES - 1) / TOK_BLOCK_BYTES;  // position n0 included
    if (nblocks > w.status_cap) return RXM_ERR_INVALID;
    if (cudaMemsetAsync(w.d_status, 0, nblocks * sizeof(uint64_t), stream) != cudaSuccess) return RXM_ERR_CUDA;
    if (cudaMemsetAsync(w.d_ticket, 0, sizeof(uint32_t), stream) != cudaSuccess) return RXM_ERR_CUDA;
    if (cudaMemsetAsync(w.d_result, 0, sizeof(unsigned long long), stream) != cudaSuccess) return RXM_ERR_CUDA;
    if (cudaMemsetAsync(w.d_result + 1, 0xff, sizeof(unsigned long long), stream) != cudaSuccess) return RXM_ERR_CUDA;
    tok_scan_kernel<<<unsigned(nblocks), TOK_THREADS, 0, stream>>>(d_text - pad, pad, n0, d_begin, d_end, cap, w.d_status,
                                                                  w.d_ticket, w.d_result, uint32_t(nblocks));
    *launched = 1;
    uint64_t eb = (cap + 255) / 256;
    const uint64_t ecap = uint64_t(sm_count) * 8;
    if (eb > ecap) eb = ecap;
    if (eb == 0) eb = 1;
    tok_exit_kernel<<<unsigned(eb), 256, 0, stream>>>(d_text, d_begin, d_end, cap, w.d_result);
    *launched = 2;
    return RXM_OK;
}

}  // namespace rxm
