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

template <int NC, int CAP, int DMAX>
__global__ void __launch_bounds__(128)
k2_mfa_thread_kernel(MfaView gv, uint32_t n_edges, const uint8_t *__restrict__ chars,
                     const uint64_t *__restrict__ offsets, uint64_t n, uint8_t *__restrict__ out,
                     unsigned long long *__restrict__ overflow) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t *s_edges = reinterpret_cast<uint64_t *>(smem);
    uint16_t *s_begin = reinterpret_cast<uint16_t *>(smem + size_t(n_edges) * 8);
    for (uint32_t i = threadIdx.x; i < n_edges; i += blockDim.x) s_edges[i] = gv.edges[i];
    for (uint32_t i = threadIdx.x; i <= gv.n_states; i += blockDim.x) s_begin[i] = gv.edge_begin[i];
    __syncthreads();
    MfaView v = gv;
    v.edges = s_edges;
    v.edge_begin = s_begin;
    MfaSim<NC, CAP, DMAX> sim;
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t b = offsets[i], e = offsets[i + 1];
        int r;
        if (e - b >= 0x7fffffffull) {
            r = 2;
        } else {
            Reader rd{chars + b, uint32_t(e - b), v.reversed};
            r = sim.run(v, rd);
        }
        if (r == 2) {
            atomicAdd(overflow, 1ull);
            r = 0;
        }
        out[i] = uint8_t(r);
    }
}

template <int NC, int CAP, int DMAX>
int launch_k2(const MfaView &v, uint32_t n_edges, const uint8_t *d_chars, const uint64_t *d_offsets,
              uint64_t n, uint8_t *d_out, unsigned long long *d_overflow, int sm_count,
              cudaStream_t stream) {
    const size_t smem = size_t(n_edges) * 8 + (size_t(v.n_states) + 1) * 2 + 16;
    auto kern = k2_mfa_thread_kernel<NC, CAP, DMAX>;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
        return RXM_ERR_CUDA;
    const int threads = 128;
    uint64_t blocks = (n + threads - 1) / threads;
    const uint64_t max_blocks = uint64_t(sm_count) * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    kern<<<unsigned(blocks), threads, smem, stream>>>(v, n_edges, d_chars, d_offsets, n, d_out, d_overflow);
    return RXM_OK;
}

}  // namespace

int k2_launch(const MfaView &v, uint32_t n_cells, uint32_t n_edges, const uint8_t *d_chars,
              const uint64_t *d_offsets, uint64_t n, uint8_t *d_out, unsigned long long *d_overflow,
              int sm_count, cudaStream_t stream, int *launched) {
    *launched = 0;
    bool rxm_dispatch_ok = true;
    int st = RXM_OK;
#define CALL(NC, CAP, DMAX) \
    st = launch_k2<NC, CAP, DMAX>(v, n_edges, d_chars, d_offsets, n, d_out, d_overflow, sm_count, stream)
    RXM_MFA_DISPATCH(n_cells, v.n_states, CALL);
#undef CALL
    if (!rxm_dispatch_ok) return RXM_ERR_UNSUPPORTED;
    if (st == RXM_OK) *launched = 1;
    return st;
}

}  // namespace rxm
