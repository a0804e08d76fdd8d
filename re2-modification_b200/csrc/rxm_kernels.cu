// sm_100a kernels of the batch matcher.  See DESIGN.md for the data layout and
// the roofline each kernel is measured against.
#include "rxm_kernels.cuh"

#include "rxm_mfa_dispatch.hpp"

namespace rxm {

// =====================================================================================
// K1 -- determinised memory-free automaton (replaces Automata::match,
//       automata.cpp:177-210, for whole batches)
// =====================================================================================

int k1_build_tables(const DfaPlan &p, K1Tables &kt, std::vector<uint8_t> &table,
                    std::vector<uint8_t> &accept, std::string *err) {
    kt = K1Tables();
    kt.n_states = p.n_states;
    kt.n_classes = p.n_classes;
    kt.start = p.start;
    kt.reversed = p.reversed;
    accept.assign(p.accept.begin(), p.accept.end());
    while (accept.size() % 16) accept.push_back(0);
    kt.accept_bytes = uint32_t(accept.size());
    if (p.n_states <= 256) {
        uint32_t l = 4;
        while ((1u << l) < p.n_states) l++;
        kt.mode = K1_DIRECT;
        kt.log2sp = l;
        const uint32_t sp = 1u << l;
        table.assign(size_t(256) * sp, 0);
        for (uint32_t b = 0; b < 256; b++)
            for (uint32_t s = 0; s < p.n_states; s++)
                table[size_t(b) * sp + s] = uint8_t(p.trans[size_t(p.byte_class[b]) * p.n_states + s]);
    } else {
        kt.mode = K1_CLASSED;
        const size_t bytes = 256 + 2 * size_t(p.n_classes) * p.n_states;
        if (bytes + kt.accept_bytes > 200 * 1024) {
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

__device__ __forceinline__ void smem_fill(uint8_t *dst, const uint8_t *__restrict__ src, uint32_t bytes) {
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
    for (uint32_t i = threadIdx.x; i < bytes / 16; i += blockDim.x) d4[i] = s4[i];
}

// One DFA step per input byte: q = T[byte][q].  Direct layout, SP = 1 << L.
template <int L>
struct DirectStep {
    const uint8_t *T;
    __device__ __forceinline__ uint32_t operator()(uint32_t q, uint32_t byte) const {
        return T[(byte << L) + q];
    }
};
struct ClassedStep {
    const uint8_t *cmap;
    const uint16_t *trans;
    uint32_t n_states;
    __device__ __forceinline__ uint32_t operator()(uint32_t q, uint32_t byte) const {
        return trans[uint32_t(cmap[byte]) * n_states + q];
    }
};

template <class Step>
__device__ __forceinline__ uint32_t step_word_fwd(const Step &st, uint32_t q, uint32_t w) {
    q = st(q, w & 0xffu);
    q = st(q, (w >> 8) & 0xffu);
    q = st(q, (w >> 16) & 0xffu);
    q = st(q, w >> 24);
    return q;
}
template <class Step>
__device__ __forceinline__ uint32_t step_word_rev(const Step &st, uint32_t q, uint32_t w) {
    q = st(q, w >> 24);
    q = st(q, (w >> 16) & 0xffu);
    q = st(q, (w >> 8) & 0xffu);
    q = st(q, w & 0xffu);
    return q;
}

// Scan one string; returns the final DFA state (0 = dead, absorbing).
template <bool REV, class Step>
__device__ __forceinline__ uint32_t scan_string(const Step &st, uint32_t q, const uint8_t *__restrict__ p,
                                                const uint8_t *__restrict__ end) {
    if (!REV) {
        while (p < end && (reinterpret_cast<uintptr_t>(p) & 15)) q = st(q, *p++);
        while (end - p >= 16 && q) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
            q = step_word_fwd(st, q, v.x);
            q = step_word_fwd(st, q, v.y);
            q = step_word_fwd(st, q, v.z);
            q = step_word_fwd(st, q, v.w);
            p += 16;
        }
        if (q)
            while (p < end) q = st(q, *p++);
    } else {
        while (p < end && (reinterpret_cast<uintptr_t>(end) & 15)) q = st(q, *--end);
        while (end - p >= 16 && q) {
            end -= 16;
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(end));
            q = step_word_rev(st, q, v.w);
            q = step_word_rev(st, q, v.z);
            q = step_word_rev(st, q, v.y);
            q = step_word_rev(st, q, v.x);
        }
        if (q)
            while (p < end) q = st(q, *--end);
    }
    return q;
}

template <bool REV, int L>
__global__ void __launch_bounds__(256)
k1_dfa_direct_kernel(const uint8_t *__restrict__ chars, const uint64_t *__restrict__ offsets, uint64_t n,
                     uint8_t *__restrict__ out, const uint8_t *__restrict__ g_table,
                     const uint8_t *__restrict__ g_accept, uint32_t accept_bytes, uint32_t start) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *T = smem;
    uint8_t *acc = smem + (256u << L);
    smem_fill(T, g_table, 256u << L);
    smem_fill(acc, g_accept, accept_bytes);
    __syncthreads();
    const DirectStep<L> st{T};
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t b = offsets[i], e = offsets[i + 1];
        const uint32_t q = scan_string<REV>(st, start, chars + b, chars + e);
        out[i] = acc[q];
    }
}

template <bool REV>
__global__ void __launch_bounds__(256)
k1_dfa_classed_kernel(const uint8_t *__restrict__ chars, const uint64_t *__restrict__ offsets, uint64_t n,
                      uint8_t *__restrict__ out, const uint8_t *__restrict__ g_table, uint32_t table_bytes,
                      const uint8_t *__restrict__ g_accept, uint32_t accept_bytes, uint32_t n_states,
                      uint32_t start) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *tab = smem;
    uint8_t *acc = smem + table_bytes;
    smem_fill(tab, g_table, table_bytes);
    smem_fill(acc, g_accept, accept_bytes);
    __syncthreads();
    const ClassedStep st{tab, reinterpret_cast<const uint16_t *>(tab + 256), n_states};
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t b = offsets[i], e = offsets[i + 1];
        const uint32_t q = scan_string<REV>(st, start, chars + b, chars + e);
        out[i] = acc[q];
    }
}

template <bool REV, int L>
int launch_direct(const K1Tables &kt, const uint8_t *d_table, const uint8_t *d_accept,
                  const uint8_t *d_chars, const uint64_t *d_offsets, uint64_t n, uint8_t *d_out,
                  int sm_count, cudaStream_t stream) {
    const size_t smem = (256u << L) + kt.accept_bytes;
    auto kern = k1_dfa_direct_kernel<REV, L>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return RXM_ERR_CUDA;
    }
    const int threads = 256;
    uint64_t blocks = (n + threads - 1) / threads;
    const uint64_t max_blocks = uint64_t(sm_count) * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    kern<<<unsigned(blocks), threads, smem, stream>>>(d_chars, d_offsets, n, d_out, d_table, d_accept,
                                                      kt.accept_bytes, kt.start);
    return RXM_OK;
}

template <bool REV>
int launch_direct_l(const K1Tables &kt, const uint8_t *d_table, const uint8_t *d_accept,
                    const uint8_t *d_chars, const uint64_t *d_offsets, uint64_t n, uint8_t *d_out,
                    int sm_count, cudaStream_t stream) {
    switch (kt.log2sp) {
        case 4: return launch_direct<REV, 4>(kt, d_table, d_accept, d_chars, d_offsets, n, d_out, sm_count, stream);
        case 5: return launch_direct<REV, 5>(kt, d_table, d_accept, d_chars, d_offsets, n, d_out, sm_count, stream);
        case 6: return launch_direct<REV, 6>(kt, d_table, d_accept, d_chars, d_offsets, n, d_out, sm_count, stream);
        case 7: return launch_direct<REV, 7>(kt, d_table, d_accept, d_chars, d_offsets, n, d_out, sm_count, stream);
        case 8: return launch_direct<REV, 8>(kt, d_table, d_accept, d_chars, d_offsets, n, d_out, sm_count, stream);
        default: return RXM_ERR_INVALID;
    }
}

}  // namespace

int k1_launch(const K1Tables &kt, const uint8_t *d_table, const uint8_t *d_accept,
              const uint8_t *d_chars, const uint64_t *d_offsets, uint64_t n, uint8_t *d_out,
              int sm_count, cudaStream_t stream, int *launched) {
    *launched = 0;
    int st;
    if (kt.mode == K1_DIRECT) {
        st = kt.reversed ? launch_direct_l<true>(kt, d_table, d_accept, d_chars, d_offsets, n, d_out, sm_count, stream)
                         : launch_direct_l<false>(kt, d_table, d_accept, d_chars, d_offsets, n, d_out, sm_count, stream);
    } else {
        const size_t smem = size_t(kt.table_bytes) + kt.accept_bytes;
        const int threads = 256;
        uint64_t blocks = (n + threads - 1) / threads;
        const uint64_t max_blocks = uint64_t(sm_count) * 8;
        if (blocks > max_blocks) blocks = max_blocks;
        if (kt.reversed) {
            auto kern = k1_dfa_classed_kernel<true>;
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
                return RXM_ERR_CUDA;
            kern<<<unsigned(blocks), threads, smem, stream>>>(d_chars, d_offsets, n, d_out, d_table, kt.table_bytes,
                                                              d_accept, kt.accept_bytes, kt.n_states, kt.start);
        } else {
            auto kern = k1_dfa_classed_kernel<false>;
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
                return RXM_ERR_CUDA;
            kern<<<unsigned(blocks), threads, smem, stream>>>(d_chars, d_offsets, n, d_out, d_table, kt.table_bytes,
                                                              d_accept, kt.accept_bytes, kt.n_states, kt.start);
        }
        st = RXM_OK;
    }
    if (st == RXM_OK) *launched = 1;
    return st;
}

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
