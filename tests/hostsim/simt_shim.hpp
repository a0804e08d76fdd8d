// TEST INFRASTRUCTURE.  A minimal single-warp SIMT emulator: enough of the CUDA device
// vocabulary (threadIdx, full-mask votes / shuffles / reductions, __syncwarp, __syncthreads,
// atomics, dynamic shared memory) to compile a warp-cooperative kernel of this repository for
// the HOST and run it -- 32 lanes as 32 fibers on one OS thread, switched at every collective.
//
// What it checks beyond results:
//   * every collective is called with the FULL mask and reached by all 32 lanes at the SAME call
//     site (source line) -- a lane that arrives somewhere else, or returns while the others wait,
//     is reported as a divergent collective / deadlock instead of hanging a GPU;
//   * a watchdog on the number of collectives per launch turns an endless loop into a report of
//     where every lane stands.
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
constexpr size_t kStack = 256 * 1024;

struct State {
    void *sp[kLanes];        // saved stack pointers of the fibers
    void *main_sp;
    bool done[kLanes];
    int ndone;
    int site[kLanes];        // call site of the collective a lane waits in (0: none)
    uint64_t slot[2][kLanes];
    unsigned gen;            // completed collectives
    int arrived;
    int cur;                 // running lane
    uint64_t collectives, limit;
    uint64_t rng;            // 0: lanes run round-robin; else the order between collectives is shuffled
    int failed;              // 0 ok, 1 divergent collective, 2 deadlock, 3 watchdog
    char msg[512];
    uint8_t *smem;
    Idx tid[kLanes];
    Idx bdim, bidx;
    std::function<void()> body;
    std::vector<uint8_t> stacks;
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
        int n = snprintf(s.msg, sizeof s.msg, "%s after %llu collectives; lane:site =", what,
                         (unsigned long long)s.collectives);
        for (int l = 0; l < kLanes && n < int(sizeof s.msg) - 12; l++)
            n += snprintf(s.msg + n, sizeof s.msg - n, " %d:%d%s", l, s.site[l], s.done[l] ? "(done)" : "");
    }
    to_main();  // never resumed
}

// Hand the processor to another lane that has not returned: the next one, or (rng != 0) a random
// one -- between two collectives the lanes of a real warp run in no particular order, and code that
// needs one (a read that must precede another lane's write, without a __syncwarp) should fail here.
inline void yield_next() {
    State &s = S();
    const int me = s.cur;
    int first = 1;
    if (s.rng) {
        s.rng ^= s.rng << 13;
        s.rng ^= s.rng >> 7;
        s.rng ^= s.rng << 17;
        first = 1 + int(s.rng % (kLanes - 1));
    }
    for (int k = 0; k < kLanes - 1; k++) {
        const int nx = (me + 1 + (first - 1 + k) % (kLanes - 1)) % kLanes;
        if (!s.done[nx]) {
            s.cur = nx;
            simt_switch(&s.sp[me], s.sp[nx]);
            return;
        }
    }
    fail(2, "deadlock: the other lanes returned while this one waits in a collective");
}

// All 32 lanes meet here.  Returns the parity of the slot buffer that now holds everybody's value.
inline unsigned meet(unsigned mask, int site, uint64_t value) {
    State &s = S();
    const int me = s.cur;
    if (mask != 0xffffffffu) fail(1, "collective with a partial mask");
    const unsigned gen = s.gen;
    s.slot[gen & 1][me] = value;
    s.site[me] = site;
    if (++s.arrived == kLanes) {
        for (int l = 0; l < kLanes; l++)
            if (s.site[l] != site) fail(1, "divergent collective: lanes met at different call sites");
        s.arrived = 0;
        s.gen = gen + 1;
        if (++s.collectives > s.limit) fail(3, "watchdog: collective limit exceeded (endless loop?)");
        if (s.rng) yield_next();  // the last lane to arrive is not always the first to go on
    } else {
        for (int l = 0; l < kLanes; l++)
            if (s.done[l]) fail(2, "deadlock: a lane returned while others wait in a collective");
        while (s.gen == gen) {
            if (s.arrived + s.ndone == kLanes) fail(2, "deadlock: every lane that has not returned waits in a collective");
            yield_next();
        }
    }
    s.site[me] = 0;
    return gen & 1;
}

inline void fiber_entry() {
    State &s = S();
    s.body();
    s.done[s.cur] = true;
    s.ndone++;
    // a lane that returns while another waits is found by that lane; otherwise run the rest
    for (int k = 1; k < kLanes; k++) {
        const int nx = (s.cur + k) % kLanes;
        if (!s.done[nx]) {
            const int me = s.cur;
            s.cur = nx;
            simt_switch(&s.sp[me], s.sp[nx]);
        }
    }
    to_main();
    abort();
}

// Run `body` as one block of 32 threads.  Returns 0, or the failure code with *msg set.
inline int launch_warp(const std::function<void()> &body, size_t smem_bytes, uint64_t limit, const char **msg,
                       uint64_t seed = 0) {
    State &s = S();
    s.body = body;
    s.stacks.assign(kStack * kLanes, 0);
    std::vector<uint8_t> smem(smem_bytes + 64, 0xcd);  // dirty: the kernel must initialise what it reads
    s.smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem.data()) + 15) & ~uintptr_t(15));
    s.gen = 0;
    s.arrived = 0;
    s.ndone = 0;
    s.collectives = 0;
    s.limit = limit;
    s.rng = seed ? seed * 0x9e3779b97f4a7c15ull + 1 : 0;
    s.failed = 0;
    s.msg[0] = 0;
    s.bdim = Idx{kLanes, 1, 1};
    s.bidx = Idx{0, 0, 0};
    for (int l = 0; l < kLanes; l++) {
        s.done[l] = false;
        s.site[l] = 0;
        s.tid[l] = Idx{unsigned(l), 0, 0};
        // initial frame: six callee-saved registers, then the entry as return address; the stack
        // pointer is 16n + 8 when the entry starts, as after a call
        uintptr_t top = reinterpret_cast<uintptr_t>(s.stacks.data() + kStack * (l + 1));
        top &= ~uintptr_t(15);
        void **p = reinterpret_cast<void **>(top);
        *--p = nullptr;                                   // fake return address of the entry (keeps alignment)
        *--p = reinterpret_cast<void *>(&fiber_entry);    // ret target of the first switch
        for (int r = 0; r < 6; r++) *--p = nullptr;
        s.sp[l] = p;
    }
    s.cur = 0;
    simt_switch(&s.main_sp, s.sp[0]);
    if (msg) *msg = s.msg;
    return s.failed;
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

inline unsigned ballot(unsigned mask, bool p, int site) {
    State &s = S();
    const unsigned par = meet(mask, site, p ? 1 : 0);
    unsigned r = 0;
    for (int l = 0; l < kLanes; l++) r |= unsigned(s.slot[par][l] & 1) << l;
    return r;
}
template <typename T>
inline T shfl(unsigned mask, T v, int src, int width, int site) {
    State &s = S();
    const int me = s.cur;
    const unsigned par = meet(mask, site, to_bits(v));
    const int base = me & ~(width - 1);
    return from_bits<T>(s.slot[par][base + (src & (width - 1))]);
}
template <typename T>
inline T shfl_up(unsigned mask, T v, unsigned d, int width, int site) {
    State &s = S();
    const int me = s.cur;
    const unsigned par = meet(mask, site, to_bits(v));
    const int rel = me & (width - 1);
    return rel >= int(d) ? from_bits<T>(s.slot[par][me - int(d)]) : v;
}
template <typename T>
inline T shfl_xor(unsigned mask, T v, int lanemask, int width, int site) {
    State &s = S();
    const int me = s.cur;
    const unsigned par = meet(mask, site, to_bits(v));
    const int other = me ^ lanemask;
    return (other & ~(width - 1)) == (me & ~(width - 1)) ? from_bits<T>(s.slot[par][other]) : v;
}
inline unsigned reduce_min(unsigned mask, unsigned v, int site) {
    State &s = S();
    const unsigned par = meet(mask, site, v);
    unsigned r = 0xffffffffu;
    for (int l = 0; l < kLanes; l++) r = s.slot[par][l] < r ? unsigned(s.slot[par][l]) : r;
    return r;
}
inline unsigned reduce_max(unsigned mask, unsigned v, int site) {
    State &s = S();
    const unsigned par = meet(mask, site, v);
    unsigned r = 0;
    for (int l = 0; l < kLanes; l++) r = s.slot[par][l] > r ? unsigned(s.slot[par][l]) : r;
    return r;
}

}  // namespace simt

// ---- the device vocabulary ----
#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __shared__
#define threadIdx (simt::S().tid[simt::S().cur])
#define blockDim (simt::S().bdim)
#define blockIdx (simt::S().bidx)

#define __syncthreads() ((void)simt::meet(0xffffffffu, __LINE__, 0))
#define __syncwarp(m) ((void)simt::meet((m), __LINE__, 0))
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

struct alignas(16) uint4 {
    unsigned x, y, z, w;
};
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz(unsigned(x)) : 32; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
    return unsigned(((uint64_t(hi) << 32) | lo) >> (shift & 31u));
}
template <typename T>
static inline T __ldg(const T *p) {
    return *p;
}
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    const unsigned long long old = *p;
    *p = old + v;
    return old;
}
static inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) {
    const unsigned long long old = *p;
    if (v < old) *p = v;
    return old;
}
template <typename T>
static inline T min(T a, T b) {
    return b < a ? b : a;
}

#endif
