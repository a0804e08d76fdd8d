// K2 -- MFA, one thread per string, sm_100a.
// Replaces MFA::match (mfa.cpp:215-236) for whole batches; the simulation itself is
// rxm_mfa_core.cuh.
#include "rxm_kernels.cuh"

#include "rxm_mfa_dispatch.hpp"

namespace rxm {

// =====================================================================================
// K2 -- MFA, one thread per string (replaces MFA::match, mfa.cpp:215-236)
// =====================================================================================
namespace {

// The per-string simulation state (MfaSim: two frontier buffers + the recursion stack)
// lives in SHARED memory, one object per thread at an odd word stride so that the 32
// lanes of a warp touching the same member hit 32 different banks.  (In local memory
// the same state is ~0.7-9 KB per thread and falls out of L1 into L2.)
template <int NC, int CAP, int DMAX>
struct K2Geom {
    typedef MfaSim<NC, CAP, DMAX> sim_t;
    static constexpr uint32_t words = (uint32_t(sizeof(sim_t)) + 3u) / 4u;
    static constexpr uint32_t stride_words = words | 1u;  // odd
    static constexpr uint32_t stride_bytes = stride_words * 4u;
};

constexpr int K2_THREADS = 128;

// LOCAL = true: the state does not fit shared memory (large automata, e.g. the 77-state
// reversed example 8): it lives in per-thread local memory instead.
template <int NC, int CAP, int DMAX, bool LOCAL>
__global__ void __launch_bounds__(K2_THREADS)
k2_mfa_thread_kernel(MfaView gv, uint32_t n_edges, const uint8_t *__restrict__ chars,
                     const Spans sp, uint64_t n, uint8_t *__restrict__ out,
                     unsigned long long *__restrict__ overflow, unsigned long long *__restrict__ next_string,
                     uint32_t threads_used) {
    typedef K2Geom<NC, CAP, DMAX> G;
    RXM_DYN_SMEM(smem);
    uint64_t *s_edges = reinterpret_cast<uint64_t *>(smem);
    uint16_t *s_begin = reinterpret_cast<uint16_t *>(smem + size_t(n_edges) * 8);
    const size_t tab_bytes = (size_t(n_edges) * 8 + (size_t(gv.n_states) + 1) * 2 + 15) & ~size_t(15);
    for (uint32_t i = threadIdx.x; i < n_edges; i += blockDim.x) s_edges[i] = gv.edges[i];
    for (uint32_t i = threadIdx.x; i <= gv.n_states; i += blockDim.x) s_begin[i] = gv.edge_begin[i];
    __syncthreads();
    if (threadIdx.x >= threads_used) return;
    MfaView v = gv;
    v.edges = s_edges;
    v.edge_begin = s_begin;
    typename G::sim_t local_sim;
    typename G::sim_t *sim =
        LOCAL ? &local_sim
              : reinterpret_cast<typename G::sim_t *>(smem + tab_bytes + size_t(threadIdx.x) * G::stride_bytes);
    // strings are handed out one at a time: their cost varies by orders of magnitude
    for (;;) {
        const uint64_t i = atomicAdd(next_string, 1ull);
        if (i >= n) break;
        const uint64_t b = sp.begin[i], e = sp.end[i];
        int r;
        if (e - b >= 0x7fffffffull) {
            r = 2;
        } else {
            Reader rd{chars + b, uint32_t(e - b), v.reversed};
            r = sim->run(v, rd);
        }
        if (r == 2) {
            atomicAdd(overflow, 1ull);
            r = 0;
        }
        out[i] = uint8_t(r);
    }
}

template <int NC, int CAP, int DMAX>
int launch_k2(const MfaView &v, uint32_t n_edges, const uint8_t *d_chars, Spans spans,
              uint64_t n, uint8_t *d_out, unsigned long long *d_overflow, unsigned long long *d_next,
              int sm_count, cudaStream_t stream) {
    typedef K2Geom<NC, CAP, DMAX> G;
    const size_t tab_bytes = (size_t(n_edges) * 8 + (size_t(v.n_states) + 1) * 2 + 15) & ~size_t(15);
    // as many threads per block as the simulation state allows (<= K2_THREADS, whole warps)
    const size_t budget = 100 * 1024;  // two blocks per SM
    uint32_t threads = K2_THREADS;
    while (threads > 32 && tab_bytes + size_t(threads) * G::stride_bytes > budget) threads -= 32;
    size_t smem = tab_bytes + size_t(threads) * G::stride_bytes;
    const bool local = smem > budget;
    if (local) {
        threads = K2_THREADS;
        smem = tab_bytes;
    }
    auto kern = local ? k2_mfa_thread_kernel<NC, CAP, DMAX, true> : k2_mfa_thread_kernel<NC, CAP, DMAX, false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 /* one value per function: launches may come from several threads */) != cudaSuccess)
        return RXM_ERR_CUDA;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, K2_THREADS, smem) != cudaSuccess || nb <= 0)
        return RXM_ERR_CUDA;
    uint64_t blocks = uint64_t(sm_count) * nb;
    const uint64_t need = (n + threads - 1) / threads;
    if (blocks > need) blocks = need;
    if (cudaMemsetAsync(d_next, 0, sizeof(unsigned long long), stream) != cudaSuccess) return RXM_ERR_CUDA;
    RXM_LAUNCH(kern, unsigned(blocks), K2_THREADS, smem, stream, v, n_edges, d_chars, spans, n, d_out, d_overflow, d_next, threads);
    return RXM_OK;
}

}  // namespace

int k2_launch(const MfaView &v, uint32_t n_cells, uint32_t n_edges, const uint8_t *d_chars,
              Spans spans, uint64_t n, uint8_t *d_out, unsigned long long *d_overflow,
              unsigned long long *d_next, int sm_count, cudaStream_t stream, int *launched) {
    *launched = 0;
    bool rxm_dispatch_ok = true;
    int st = RXM_OK;
#define CALL(NC, CAP, DMAX) \
    st = launch_k2<NC, CAP, DMAX>(v, n_edges, d_chars, spans, n, d_out, d_overflow, d_next, sm_count, stream)
    RXM_MFA_DISPATCH(n_cells, v.n_states, CALL);
#undef CALL
    if (!rxm_dispatch_ok) return RXM_ERR_UNSUPPORTED;
    if (st == RXM_OK) *launched = 1;
    return st;
}

}  // namespace rxm
