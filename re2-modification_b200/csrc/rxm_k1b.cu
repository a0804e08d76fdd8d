// K1B -- memory-free automaton as a bit set, sm_100a.
// Replaces Automata::match (automata.cpp:177-210) for automata whose determinisation under the
// reference's exact step is too large for K1's table (rxm_plan.cpp: kMaxDfaStates).  One thread
// per string; the active set std::set<Node*> is a 128-bit mask in registers (bit = address rank,
// so ascending bits == set order, automata.cpp:122).  Two kernels:
//   k1b_mask_kernel    bit-parallel: next = OR over r in S of LS[class][r] & ~(S & bits_below(r)),
//                      follow masks LS in shared memory (rxm_plan.hpp: BitsetMasks explains why this
//                      IS the reference's step, `visited` quirk included, when no letter-edge
//                      target has an incoming epsilon edge)
//   k1b_bitset_kernel  the general form: `visited` and the next set are masks too, and
//                      Automata::evaluateState's recursion through epsilon edges
//                      (automata.cpp:108-110) is an explicit stack of (node, next edge); an edge --
//                      letter edges too -- whose target was already evaluated in this step is skipped
//                      (automata.cpp:104-107), a node is marked evaluated only after its edge loop (:116)
// Bound: instruction issue (per input byte ~25 instructions per member of the active set).
#include "rxm_kernels.cuh"
#include "rxm_nfa_core.cuh"

namespace rxm {

namespace {

constexpr int K1B_THREADS = 128;
static_assert(kNfaBitsDepth == int(kBitsetMaxDepth), "planner limit == kernel stack");

__global__ void __launch_bounds__(K1B_THREADS)
k1b_bitset_kernel(const uint16_t *__restrict__ g_eb, const uint32_t *__restrict__ g_ed, uint32_t n_states,
                  uint32_t n_edges, uint32_t start, uint32_t finish, uint32_t reversed,
                  const uint8_t *__restrict__ chars, const Spans sp, const K1Rec *__restrict__ recs, uint64_t n,
                  uint8_t *__restrict__ out,
                  unsigned long long *__restrict__ overflow, unsigned long long *__restrict__ next_string) {
    RXM_DYN_SMEM(smem);
    uint32_t *ed = reinterpret_cast<uint32_t *>(smem);
    uint16_t *eb = reinterpret_cast<uint16_t *>(smem + size_t(n_edges) * 4);
    for (uint32_t i = threadIdx.x; i < n_edges; i += blockDim.x) ed[i] = g_ed[i];
    for (uint32_t i = threadIdx.x; i <= n_states; i += blockDim.x) eb[i] = g_eb[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(next_string, 32ull);
        base = __shfl_sync(0xffffffffu, base, 0);
        uint64_t i = base + lane;
        if (recs) {
            // the tile sort's order (see rxm_k3.cu): 32 consecutive tickets are one group of one tile,
            // i.e. 32 strings of nearly equal length -- the lanes of this warp finish together
            const uint64_t ntiles = (n + K1_TILE_STRINGS - 1) / K1_TILE_STRINGS;
            const uint64_t g = base / (ntiles * 32u), tl = (base - g * (ntiles * 32u)) >> 5;
            if (g >= K1_TILE_STRINGS / 32u) break;
            const uint64_t pos = tl * K1_TILE_STRINGS + g * 32u + lane;
            if (pos >= min(n, (tl + 1u) * K1_TILE_STRINGS)) continue;
            i = recs[pos].idx;
        } else {
            if (base >= n) break;
            if (i >= n) continue;
        }
        const uint64_t b = sp.begin[i], e = sp.end[i];
        bool ok = e - b < 0x7fffffffull;
        Bits128 S{0, 0}, N;
        S.set(start);  // automata.cpp:178-179
        const uint8_t *s = chars + b;
        const uint32_t len = ok ? uint32_t(e - b) : 0u;
        for (uint32_t k = 0; k < len && ok; k++) {  // :181-200
            const int letter = reversed ? s[len - 1u - k] : s[k];
            ok = nfa_bits_step(eb, ed, finish, S, letter, N);
            S = N;
            if (S.empty()) break;  // :186-188
        }
        if (ok) ok = nfa_bits_step(eb, ed, finish, S, -1, N);  // :201-202
        if (!ok) {
            atomicAdd(overflow, 1ull);
            out[i] = 0;
        } else {
            out[i] = N.test(finish) ? 1 : 0;  // :204-209
        }
    }
}

__device__ __forceinline__ uint4 k1b_ld128(const uint8_t *p) {  // p is 16-byte aligned
#if defined(__CUDA_ARCH__)
    return __ldg(reinterpret_cast<const uint4 *>(p));
#else
#ifdef RXM_SIMT_HOST  // the emulator reports what the device would fault on
    if (reinterpret_cast<uintptr_t>(p) & 15u) simt::fail(4, "misaligned 16-byte global load");
#endif
    return *reinterpret_cast<const uint4 *>(p);
#endif
}
__device__ __forceinline__ uint32_t k1b_vec_byte(const uint4 &v, uint32_t j) {  // byte j (memory order) of a vector
    const uint32_t w = (j & 8u) ? ((j & 4u) ? v.w : v.z) : ((j & 4u) ? v.y : v.x);
    return (w >> ((j & 3u) * 8u)) & 0xffu;
}

// Bit-parallel form (rxm_plan.hpp: BitsetMasks): per input byte one class lookup and, per root of
// the active set, one 16-byte follow-mask row from shared memory -- no edge walk, no stack.
__global__ void __launch_bounds__(K1B_THREADS)
k1b_mask_kernel(const uint64_t *__restrict__ g_ls, const uint8_t *__restrict__ g_class, uint32_t n_states,
                uint32_t n_classes, uint32_t start, uint64_t acc_lo, uint64_t acc_hi, uint32_t reversed,
                const uint8_t *__restrict__ chars, const Spans sp, const K1Rec *__restrict__ recs, uint64_t n,
                uint8_t *__restrict__ out, unsigned long long *__restrict__ overflow,
                unsigned long long *__restrict__ next_string) {
    RXM_DYN_SMEM(smem);
    uint64_t *ls = reinterpret_cast<uint64_t *>(smem);
    uint8_t *bc = smem + size_t(n_classes) * n_states * 16;
    for (uint32_t i = threadIdx.x; i < n_classes * n_states * 2; i += blockDim.x) ls[i] = g_ls[i];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) bc[i] = g_class[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(next_string, 32ull);
        base = __shfl_sync(0xffffffffu, base, 0);
        uint64_t i = base + lane;
        if (recs) {  // the tile sort's order: 32 strings of nearly equal length per warp
            const uint64_t ntiles = (n + K1_TILE_STRINGS - 1) / K1_TILE_STRINGS;
            const uint64_t g = base / (ntiles * 32u), tl = (base - g * (ntiles * 32u)) >> 5;
            if (g >= K1_TILE_STRINGS / 32u) break;
            const uint64_t pos = tl * K1_TILE_STRINGS + g * 32u + lane;
            if (pos >= min(n, (tl + 1u) * K1_TILE_STRINGS)) continue;
            i = recs[pos].idx;
        } else {
            if (base >= n) break;
            if (i >= n) continue;
        }
        const uint64_t b = sp.begin[i], e = sp.end[i];
        if (e - b >= 0x7fffffffull) {
            atomicAdd(overflow, 1ull);
            out[i] = 0;
            continue;
        }
        const uint8_t *s = chars + b;
        const uint32_t len = uint32_t(e - b);
        Bits128 S{0, 0};
        S.set(start);  // automata.cpp:178-179
        // :181-200, break on the empty set (:186-188).  The string is read in aligned 16-byte vectors (only those that
        // hold a byte of it), forward from its first byte or downwards from its last.
        for (uint32_t k = 0; k < len && !S.empty();) {
            const uint8_t *at = reversed ? s + (len - 1u - k) : s + k;  // the next byte read
            const uint32_t o = uint32_t(reinterpret_cast<uintptr_t>(at) & 15u);
            const uint4 v = k1b_ld128(at - o);
            const uint32_t left = len - k;
            if (!reversed) {
                const uint32_t hi = o + left < 16u ? o + left : 16u;
                for (uint32_t j = o; j < hi && !S.empty(); j++)
                    S = nfa_mask_step(ls + size_t(bc[k1b_vec_byte(v, j)]) * n_states * 2, S);
                k += hi - o;
            } else {
                const uint32_t take = o + 1u < left ? o + 1u : left;  // bytes o, o-1, ... of the vector
                for (uint32_t j = 0; j < take && !S.empty(); j++)
                    S = nfa_mask_step(ls + size_t(bc[k1b_vec_byte(v, o - j)]) * n_states * 2, S);
                k += take;
            }
        }
        out[i] = ((S.lo & acc_lo) | (S.hi & acc_hi)) ? 1 : 0;  // :201-209
    }
}

}  // namespace

int k1b_mask_launch(const uint64_t *d_ls, const uint8_t *d_class, uint32_t n_states, uint32_t n_classes, uint32_t start,
                    uint64_t acc_lo, uint64_t acc_hi, uint32_t reversed, const uint8_t *d_chars, Spans spans,
                    const K1Rec *d_recs, uint64_t n, uint8_t *d_out, unsigned long long *d_overflow,
                    unsigned long long *d_next, int sm_count, cudaStream_t stream, int *launched) {
    *launched = 0;
    const size_t smem = size_t(n_classes) * n_states * 16 + 256;
    if (smem > 96 * 1024) return RXM_ERR_UNSUPPORTED;
    if (cudaFuncSetAttribute(k1b_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 /* one value per function: launches may come from several threads */) != cudaSuccess)
        return RXM_ERR_CUDA;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k1b_mask_kernel, K1B_THREADS, smem) != cudaSuccess || nb <= 0)
        return RXM_ERR_CUDA;
    uint64_t blocks = uint64_t(sm_count) * nb;
    const uint64_t need = (n + K1B_THREADS - 1) / K1B_THREADS;
    if (blocks > need) blocks = need;
    if (cudaMemsetAsync(d_next, 0, sizeof(unsigned long long), stream) != cudaSuccess) return RXM_ERR_CUDA;
    RXM_LAUNCH(k1b_mask_kernel, unsigned(blocks), K1B_THREADS, smem, stream, d_ls, d_class, n_states, n_classes, start, acc_lo, acc_hi, reversed, d_chars, spans, d_recs, n, d_out, d_overflow, d_next);
    *launched = 1;
    return RXM_OK;
}

void k1b_build_tables(const rxm_tables &t, std::vector<uint16_t> &eb, std::vector<uint32_t> &ed) {
    eb.resize(t.n_states + 1);
    ed.resize(t.n_edges);
    for (uint32_t q = 0; q <= t.n_states; q++) eb[q] = uint16_t(t.edge_begin[q]);
    for (uint32_t e = 0; e < t.n_edges; e++) ed[e] = nfa_pack_edge(t.edge_kind[e], t.edge_sym[e], t.edge_to[e]);
}

int k1b_launch(const uint16_t *d_eb, const uint32_t *d_ed, uint32_t n_states, uint32_t n_edges, uint32_t start,
               uint32_t finish, uint32_t reversed, const uint8_t *d_chars, Spans spans, const K1Rec *d_recs, uint64_t n,
               uint8_t *d_out, unsigned long long *d_overflow, unsigned long long *d_next, int sm_count,
               cudaStream_t stream, int *launched) {
    *launched = 0;
    const size_t smem = ((size_t(n_edges) * 4 + (size_t(n_states) + 1) * 2) + 15) & ~size_t(15);
    if (smem > 96 * 1024) return RXM_ERR_UNSUPPORTED;
    if (cudaFuncSetAttribute(k1b_bitset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 /* one value per function: launches may come from several threads */) != cudaSuccess)
        return RXM_ERR_CUDA;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k1b_bitset_kernel, K1B_THREADS, smem) != cudaSuccess || nb <= 0)
        return RXM_ERR_CUDA;
    uint64_t blocks = uint64_t(sm_count) * nb;
    const uint64_t need = (n + K1B_THREADS - 1) / K1B_THREADS;
    if (blocks > need) blocks = need;
    if (cudaMemsetAsync(d_next, 0, sizeof(unsigned long long), stream) != cudaSuccess) return RXM_ERR_CUDA;
    RXM_LAUNCH(k1b_bitset_kernel, unsigned(blocks), K1B_THREADS, smem, stream, d_eb, d_ed, n_states, n_edges, start, finish, reversed, d_chars, spans, d_recs, n, d_out, d_overflow, d_next);
    *launched = 1;
    return RXM_OK;
}

}  // namespace rxm
