// TEST INFRASTRUCTURE.  A minimal SIMT emulator: enough of the CUDA vocabulary (threadIdx /
// blockIdx, full-mask votes / shuffles / reductions, __syncwarp, __syncthreads, atomics, static
// and dynamic shared memory, <<<>>> through RXM_LAUNCH, the handful of runtime calls the launch
// functions make) to compile the kernel SOURCES of this repository for the HOST and run them --
// the threads of a block as fibers on one OS thread, switched at every collective, the blocks of
// a grid one after another.
//
// What it checks beyond results:
//   * every warp collective is called with the FULL mask and reached by all 32 lanes of the warp
//     at the SAME call site (source line); every __syncthreads by all threads that have not
//     returned, at one call site -- a thread that arrives somewhere else, or returns while its
//     warp waits, is reported as a divergent collective / deadlock instead of hanging a GPU;
//   * a budget on the number of collectives per launch turns an endless loop into a report of
//     where the threads stand;
//   * with a seed the threads run in a SHUFFLED order between collectives (a real warp guarantees
//     none): code that needs an order without a __syncwarp / __syncthreads gives different results.
// It is only ever built into tests/hostsim/libhostsim.so; librxm.so does not contain it.
#ifndef RXM_SIMT_SHIM_HPP
#define RXM_SIMT_SHIM_HPP

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <vector>

#if !defined(__x86_64__)
#error "simt_shim.hpp switches fibers with x86-64 assembly"
#endif

namespace simt {

struct Idx {
    unsigned x, y, z;
};

constexpr int kLanes = 32;
constexpr int kMaxThreads = 1024;
constexpr size_t kStack = 192 * 1024;

struct WarpSync {
    unsigned gen;  // completed collectives of this warp
    int arrived;
    uint64_t slot[2][kLanes];
};

struct State {
    void *sp[kMaxThreads];  // saved stack pointers of the fibers
    void *main_sp;
    bool done[kMaxThreads];
    bool blocked[kMaxThreads];  // waits in a collective that is not complete yet: not scheduled
    int site[kMaxThreads];  // call site of the collective a thread waits in (0: none)
    WarpSync warp[kMaxThreads / kLanes];
    unsigned bgen;          // completed __syncthreads
    int barrived;
    int nthreads, ndone;
    int cur;                // running thread
    uint64_t collectives, limit;
    uint64_t rng;           // 0: threads run round-robin; else the order between collectives is shuffled
    int failed;             // 0 ok, 1 divergent collective, 2 deadlock, 3 budget
    char msg[512];
    uint8_t *smem;
    Idx tid[kMaxThreads];
    Idx bdim, bidx, gdim;
    std::function<void()> body;
    uint8_t *stacks;
    size_t stacks_size;
    // launch defaults for RXM_LAUNCH (set by the test driver), result of the last failed launch
    uint64_t default_limit, default_seed;
    int last_failure;
    char last_msg[512];
    uint64_t launches;
};
inline State &S() {
    static State s;
    return s;
}

extern "C" void simt_switch(void **from_sp, void *to_sp);
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size simt_switch,.-simt_switch
)");

inline void to_main() {
    State &s = S();
    simt_switch(&s.sp[s.cur], s.main_sp);
}

inline void fail(int code, const char *what) {
    State &s = S();
    if (!s.failed) {
        s.failed = code;
        int n = snprintf(s.msg, sizeof s.msg, "%s (block %u, thread %d) after %llu collectives; thread:site =", what,
                         s.bidx.x, s.cur, (unsigned long long)s.collectives);
        const int w0 = (s.cur / kLanes) * kLanes;  // the warp of the thread that noticed
        for (int l = w0; l < w0 + kLanes && l < s.nthreads && n < int(sizeof s.msg) - 16; l++)
            n += snprintf(s.msg + n, sizeof s.msg - n, " %d:%d%s", l, s.site[l], s.done[l] ? "(done)" : "");
    }
    to_main();  // never resumed
}

// Hand the processor to another thread that can run (has not returned, does not wait in an
// incomplete collective): the next one, or (rng != 0) a random one.  If none can and the caller
// waits itself, nobody will ever release it.
inline void yield_next() {
    State &s = S();
    const int me = s.cur, n = s.nthreads;
    if (n > 1) {
        int first = 0;
        if (s.rng) {
            s.rng ^= s.rng << 13;
            s.rng ^= s.rng >> 7;
            s.rng ^= s.rng << 17;
            first = int(s.rng % uint64_t(n - 1));
        }
        for (int k = 0; k < n - 1; k++) {
            const int nx = (me + 1 + (first + k) % (n - 1)) % n;
            if (!s.done[nx] && !s.blocked[nx]) {
                s.cur = nx;
                simt_switch(&s.sp[me], s.sp[nx]);
                return;
            }
        }
    }
    if (s.blocked[me]) fail(2, "deadlock: every thread that has not returned waits in a collective");
}

inline void count_collective() {
    State &s = S();
    if (++s.collectives > s.limit) fail(3, "watchdog: collective limit exceeded (endless loop?)");
}

// The 32 lanes of the calling thread's warp meet here.  Returns the parity of the slot buffer that
// now holds everybody's value.
inline unsigned meet(unsigned mask, int site, uint64_t value) {
    State &s = S();
    const int me = s.cur, w = me / kLanes, w0 = w * kLanes;
    WarpSync &W = s.warp[w];
    if (mask != 0xffffffffu) fail(1, "collective with a partial mask");
    const unsigned gen = W.gen;
    W.slot[gen & 1][me - w0] = value;
    s.site[me] = site;
    for (int l = w0; l < w0 + kLanes; l++)
        if (s.done[l]) fail(2, "deadlock: a lane returned while its warp meets in a collective");
    if (++W.arrived == kLanes) {
        for (int l = w0; l < w0 + kLanes; l++)
            if (s.site[l] != site) fail(1, "divergent collective: the lanes of a warp met at different call sites");
        W.arrived = 0;
        W.gen = gen + 1;
        for (int l = w0; l < w0 + kLanes; l++) s.blocked[l] = false;
        count_collective();
        if (s.rng) yield_next();  // the last lane to arrive is not always the first to go on
    } else {
        s.blocked[me] = true;
        while (W.gen == gen) yield_next();
    }
    s.site[me] = 0;
    return gen & 1;
}

inline void release_block_barrier() {
    State &s = S();
    s.barrived = 0;
    s.bgen++;
    for (int l = 0; l < s.nthreads; l++)
        if (s.site[l] < 0) s.blocked[l] = false;  // the threads in __syncthreads
}

// __syncthreads: every thread of the block that has not returned.
inline void sync_block(int site) {
    State &s = S();
    const int me = s.cur;
    s.site[me] = -site;  // negative: block-level
    const unsigned gen = s.bgen;
    if (++s.barrived == s.nthreads - s.ndone) {
        for (int l = 0; l < s.nthreads; l++)
            if (!s.done[l] && s.site[l] != -site) fail(1, "divergent __syncthreads: threads met at different call sites");
        release_block_barrier();
        count_collective();
        if (s.rng) yield_next();
    } else {
        s.blocked[me] = true;
        while (s.bgen == gen) yield_next();
    }
    s.site[me] = 0;
}

inline void fiber_entry() {
    State &s = S();
    s.body();
    s.done[s.cur] = true;
    s.ndone++;
    // threads that returned do not take part in __syncthreads
    if (s.barrived > 0 && s.barrived == s.nthreads - s.ndone) release_block_barrier();
    for (int k = 1; k < s.nthreads; k++) {
        const int nx = (s.cur + k) % s.nthreads;
        if (!s.done[nx]) {
            const int me = s.cur;
            s.cur = nx;
            simt_switch(&s.sp[me], s.sp[nx]);
        }
    }
    to_main();
    abort();
}

// Run `body` as ONE block of `nthreads` threads (a multiple of 32).  Returns 0, or the failure code
// with *msg set.
inline int launch_block(const std::function<void()> &body, int nthreads, uint8_t *smem, uint64_t limit, const char **msg,
                        uint64_t seed) {
    State &s = S();
    if (nthreads <= 0 || nthreads > kMaxThreads || (nthreads % kLanes) != 0) {
        snprintf(s.msg, sizeof s.msg, "block size %d is not a multiple of 32 up to %d", nthreads, kMaxThreads);
        if (msg) *msg = s.msg;
        return 1;
    }
    s.body = body;
    const size_t need = kStack * size_t(nthreads);
    if (s.stacks_size < need) {
        free(s.stacks);
        s.stacks = static_cast<uint8_t *>(malloc(need));  // pages are touched on demand
        s.stacks_size = need;
    }
    s.smem = smem;
    s.nthreads = nthreads;
    s.ndone = 0;
    s.bgen = 0;
    s.barrived = 0;
    s.collectives = 0;
    s.limit = limit;
    s.rng = seed ? seed * 0x9e3779b97f4a7c15ull + 1 : 0;
    s.failed = 0;
    s.msg[0] = 0;
    s.bdim = Idx{unsigned(nthreads), 1, 1};
    for (int w = 0; w < nthreads / kLanes; w++) {
        s.warp[w].gen = 0;
        s.warp[w].arrived = 0;
    }
    for (int l = 0; l < nthreads; l++) {
        s.done[l] = false;
        s.blocked[l] = false;
        s.site[l] = 0;
        s.tid[l] = Idx{unsigned(l), 0, 0};
        // initial frame: six callee-saved registers, then the entry as return address; the stack
        // pointer is 16n + 8 when the entry starts, as after a call
        uintptr_t top = reinterpret_cast<uintptr_t>(s.stacks + kStack * size_t(l + 1));
        top &= ~uintptr_t(15);
        void **p = reinterpret_cast<void **>(top);
        *--p = nullptr;                                 // fake return address of the entry (keeps alignment)
        *--p = reinterpret_cast<void *>(&fiber_entry);  // ret target of the first switch
        for (int r = 0; r < 6; r++) *--p = nullptr;
        s.sp[l] = p;
    }
    s.cur = 0;
    simt_switch(&s.main_sp, s.sp[0]);
    if (msg) *msg = s.msg;
    return s.failed;
}

// One warp (the emulator's self-tests use it).
inline int launch_warp(const std::function<void()> &body, size_t smem_bytes, uint64_t limit, const char **msg,
                       uint64_t seed = 0) {
    std::vector<uint8_t> smem(smem_bytes + 256, 0xcd);  // dirty: the kernel must initialise what it reads
    State &s = S();
    s.bidx = Idx{0, 0, 0};
    s.gdim = Idx{1, 1, 1};
    return launch_block(body, kLanes,
                        reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem.data()) + 127) & ~uintptr_t(127)), limit,
                        msg, seed);
}

// kernel<<<grid, block, smem>>>(...): the blocks one after another.  A failure is kept in
// last_failure / last_msg (the library's launch functions cannot see it) and later launches of
// the run are skipped.
inline void launch_grid(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()> &body) {
    State &s = S();
    s.launches++;
    if (s.last_failure) return;
    std::vector<uint8_t> smem(smem_bytes + 256, 0xcd);
    uint8_t *sm = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem.data()) + 127) & ~uintptr_t(127));
    s.gdim = Idx{grid, 1, 1};
    for (unsigned b = 0; b < grid; b++) {
        s.bidx = Idx{b, 0, 0};
        const char *msg = "";
        const uint64_t seed = s.default_seed ? s.default_seed + b : 0;
        const int rc = launch_block(body, int(block), sm, s.default_limit ? s.default_limit : 400000000ull, &msg, seed);
        if (rc) {
            s.last_failure = rc;
            snprintf(s.last_msg, sizeof s.last_msg, "%s", msg);
            return;
        }
    }
}

template <typename T>
inline T from_bits(uint64_t b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}
template <typename T>
inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8, "");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
inline int my_lane() { return S().cur % kLanes; }
inline const uint64_t *slots(unsigned par) { return S().warp[S().cur / kLanes].slot[par]; }

inline unsigned ballot(unsigned mask, bool p, int site) {
    const unsigned par = meet(mask, site, p ? 1 : 0);
    const uint64_t *sl = slots(par);
    unsigned r = 0;
    for (int l = 0; l < kLanes; l++) r |= unsigned(sl[l] & 1) << l;
    return r;
}
template <typename T>
inline T shfl(unsigned mask, T v, int src, int width, int site) {
    const int me = my_lane();
    const unsigned par = meet(mask, site, to_bits(v));
    const int base = me & ~(width - 1);
    return from_bits<T>(slots(par)[base + (src & (width - 1))]);
}
template <typename T>
inline T shfl_up(unsigned mask, T v, unsigned d, int width, int site) {
    const int me = my_lane();
    const unsigned par = meet(mask, site, to_bits(v));
    const int rel = me & (width - 1);
    return rel >= int(d) ? from_bits<T>(slots(par)[me - int(d)]) : v;
}
template <typename T>
inline T shfl_xor(unsigned mask, T v, int lanemask, int width, int site) {
    const int me = my_lane();
    const unsigned par = meet(mask, site, to_bits(v));
    const int other = me ^ lanemask;
    return (other & ~(width - 1)) == (me & ~(width - 1)) ? from_bits<T>(slots(par)[other]) : v;
}
inline unsigned reduce_min(unsigned mask, unsigned v, int site) {
    const unsigned par = meet(mask, site, v);
    const uint64_t *sl = slots(par);
    unsigned r = 0xffffffffu;
    for (int l = 0; l < kLanes; l++) r = sl[l] < r ? unsigned(sl[l]) : r;
    return r;
}
inline unsigned reduce_max(unsigned mask, unsigned v, int site) {
    const unsigned par = meet(mask, site, v);
    const uint64_t *sl = slots(par);
    unsigned r = 0;
    for (int l = 0; l < kLanes; l++) r = sl[l] > r ? unsigned(sl[l]) : r;
    return r;
}
inline unsigned reduce_add(unsigned mask, unsigned v, int site) {
    const unsigned par = meet(mask, site, v);
    const uint64_t *sl = slots(par);
    unsigned r = 0;
    for (int l = 0; l < kLanes; l++) r += unsigned(sl[l]);
    return r;
}

// 32-bit shared-memory addresses (PTX ld.shared / cp.async operands): offsets into the emulated
// block's dynamic shared memory, biased so that 0 is never a valid one.
constexpr uint32_t kSharedBias = 1u << 16;  // a multiple of every alignment the kernels ask for
inline uint8_t *shared_ptr(uint32_t addr) { return S().smem + (addr - kSharedBias); }
inline unsigned mask_or_full() { return 0xffffffffu; }  // __syncwarp() without a mask
inline unsigned mask_or_full(unsigned m) { return m; }

}  // namespace simt

// ---- the device vocabulary ----
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static  /* one block runs at a time: a function-scope static is the block's array */
#define threadIdx (simt::S().tid[simt::S().cur])
#define blockDim (simt::S().bdim)
#define blockIdx (simt::S().bidx)
#define gridDim (simt::S().gdim)

#define __syncthreads() simt::sync_block(__LINE__)
#define __syncwarp(...) ((void)simt::meet(simt::mask_or_full(__VA_ARGS__), __LINE__, 0))
#define __ballot_sync(m, p) simt::ballot((m), (p), __LINE__)
#define __any_sync(m, p) (simt::ballot((m), (p), __LINE__) != 0u)
#define __all_sync(m, p) (simt::ballot((m), (p), __LINE__) == 0xffffffffu)
#define SIMT_SHFL3(m, v, s) simt::shfl((m), (v), (s), 32, __LINE__)
#define SIMT_SHFL4(m, v, s, w) simt::shfl((m), (v), (s), (w), __LINE__)
#define SIMT_PICK(_1, _2, _3, _4, NAME, ...) NAME
#define __shfl_sync(...) SIMT_PICK(__VA_ARGS__, SIMT_SHFL4, SIMT_SHFL3)(__VA_ARGS__)
#define SIMT_UP3(m, v, d) simt::shfl_up((m), (v), (d), 32, __LINE__)
#define SIMT_UP4(m, v, d, w) simt::shfl_up((m), (v), (d), (w), __LINE__)
#define __shfl_up_sync(...) SIMT_PICK(__VA_ARGS__, SIMT_UP4, SIMT_UP3)(__VA_ARGS__)
#define SIMT_XOR3(m, v, d) simt::shfl_xor((m), (v), (d), 32, __LINE__)
#define SIMT_XOR4(m, v, d, w) simt::shfl_xor((m), (v), (d), (w), __LINE__)
#define __shfl_xor_sync(...) SIMT_PICK(__VA_ARGS__, SIMT_XOR4, SIMT_XOR3)(__VA_ARGS__)
#define __reduce_min_sync(m, v) simt::reduce_min((m), (v), __LINE__)
#define __reduce_max_sync(m, v) simt::reduce_max((m), (v), __LINE__)
#define __reduce_add_sync(m, v) simt::reduce_add((m), (v), __LINE__)

struct alignas(16) uint4 {
    unsigned x, y, z, w;
};
struct alignas(8) uint2 {
    unsigned x, y;
};
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clz(int x) { return x ? __builtin_clz(unsigned(x)) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline unsigned __brev(unsigned x) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
    return unsigned(((uint64_t(hi) << 32) | lo) >> (shift & 31u));
}
static inline size_t __cvta_generic_to_shared(const void *p) {
    return size_t(static_cast<const uint8_t *>(p) - simt::S().smem) + simt::kSharedBias;
}
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned sel) {
    const uint64_t v = (uint64_t(y) << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= unsigned((v >> (8 * ((sel >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
}
template <typename T>
static inline T __ldg(const T *p) {
    return *p;
}
template <typename T, typename U>
static inline T atomicAdd(T *p, U v) {
    const T old = *p;
    *p = old + T(v);
    return old;
}
template <typename T, typename U>
static inline T atomicMin(T *p, U v) {
    const T old = *p;
    if (T(v) < old) *p = T(v);
    return old;
}
template <typename T, typename U>
static inline T atomicMax(T *p, U v) {
    const T old = *p;
    if (T(v) > old) *p = T(v);
    return old;
}
template <typename T, typename U>
static inline T atomicOr(T *p, U v) {
    const T old = *p;
    *p = old | T(v);
    return old;
}
template <typename T>
static inline T min(T a, T b) {
    return b < a ? b : a;
}
template <typename T>
static inline T max(T a, T b) {
    return a < b ? b : a;
}

// ---- the runtime calls the launch functions make ----
typedef void *cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F>
static inline cudaError_t cudaFuncSetAttribute(F, int, int) {
    return cudaSuccess;
}
template <typename F>
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *nb, F, int, size_t) {
    *nb = 1;
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) {
    memset(p, v, n);
    return cudaSuccess;
}
#define RXM_LAUNCH(kern, grid, block, smem, stream, ...) \
    simt::launch_grid((grid), (block), (smem), [&]() { kern(__VA_ARGS__); })
#define RXM_DYN_SMEM(name) uint8_t *name = simt::S().smem
#define RXM_DYN_SMEM_128(name) uint8_t *name = simt::S().smem

#endif
