// Bit-set simulation of a memory-free automaton, shared by the K1B kernel (rxm_k1b.cu) and, on
// the host, by tests/hostsim.  One call of nfa_bits_step is one Automata::evaluateStates
// (automata.cpp:119-128) with evaluateState (automata.cpp:98-117) on flat tables; see rxm_k1b.cu.
#ifndef RXM_NFA_CORE_CUH
#define RXM_NFA_CORE_CUH

#include <stdint.h>

#include "../../include/rxm.h"
#include "rxm_mfa_core.cuh"  // RXM_HD

namespace rxm {

constexpr int kNfaBitsDepth = 48;  // == kBitsetMaxDepth (rxm_plan.hpp): explicit recursion stack

RXM_HD uint32_t nfa_ctz64(uint64_t w) {
#if defined(__CUDA_ARCH__)
    return uint32_t(__ffsll(static_cast<long long>(w)) - 1);
#else
    return uint32_t(__builtin_ctzll(w));
#endif
}

struct Bits128 {
    uint64_t lo, hi;
    RXM_HD bool test(uint32_t q) const { return ((q < 64 ? lo >> q : hi >> (q - 64)) & 1ull) != 0; }
    RXM_HD void set(uint32_t q) {
        if (q < 64) lo |= 1ull << q;
        else hi |= 1ull << (q - 64);
    }
    RXM_HD bool empty() const { return (lo | hi) == 0; }
};

// packed edge: kind:2 | sym:8 | to:16
RXM_HD uint32_t nfa_pack_edge(uint32_t kind, uint32_t sym, uint32_t to) {
    return kind | (sym << 2) | (to << 10);
}

// one Automata::evaluateStates call (automata.cpp:119-128); letter < 0 is the final pass
RXM_HD bool nfa_bits_step(const uint16_t *eb, const uint32_t *ed, uint32_t finish, Bits128 S, int letter,
                                         Bits128 &N) {
    Bits128 visited{0, 0};
    N = Bits128{0, 0};
    uint32_t stack[kNfaBitsDepth];  // node:8 | next edge:24
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
        uint64_t w = half ? S.hi : S.lo;
        while (w) {
            const uint32_t q = nfa_ctz64(w) + 64u * uint32_t(half);
            w &= w - 1;
            if (visited.test(q)) continue;  // :123
            int sp = 0;
            stack[0] = q | (uint32_t(eb[q]) << 8);
            while (sp >= 0) {
                const uint32_t u = stack[sp] & 0xffu, e = stack[sp] >> 8;
                if (letter < 0 && u == finish) {  // :100-102
                    N.set(u);
                    visited.set(u);  // :116
                    sp--;
                    continue;
                }
                if (e == eb[u + 1]) {
                    visited.set(u);  // :116, after the loop
                    sp--;
                    continue;
                }
                stack[sp] = u | ((e + 1) << 8);
                const uint32_t x = ed[e], kind = x & 3u, to = x >> 10;
                if (visited.test(to)) continue;  // :105-107 -- letter edges too
                if (kind == RXM_EDGE_EPS) {      // :108-110
                    if (sp + 1 >= kNfaBitsDepth) return false;
                    stack[++sp] = to | (uint32_t(eb[to]) << 8);
                } else if (letter >= 0 && (kind == RXM_EDGE_ANY || (kind == RXM_EDGE_LIT && int((x >> 2) & 0xffu) == letter))) {
                    N.set(to);  // :111-113
                }
            }
        }
    }
    return true;
}

// The same step from follow masks (rxm_plan.hpp: BitsetMasks): ls = LS[class] = n_states x 2 words.
RXM_HD Bits128 nfa_mask_step(const uint64_t *ls, Bits128 S) {
    Bits128 N{0, 0};
    uint64_t w = S.lo;
    while (w) {  // roots 0..63: the filter is S.lo below r
        const uint32_t r = nfa_ctz64(w);
        const uint64_t below = S.lo & ((1ull << r) - 1ull);
        w &= w - 1;
        N.lo |= ls[2 * r] & ~below;
        N.hi |= ls[2 * r + 1];
    }
    w = S.hi;
    while (w) {  // roots 64..127: all of S.lo lies below
        const uint32_t r = nfa_ctz64(w);
        const uint64_t below = S.hi & ((1ull << r) - 1ull);
        w &= w - 1;
        N.lo |= ls[2 * (r + 64)] & ~S.lo;
        N.hi |= ls[2 * (r + 64) + 1] & ~below;
    }
    return N;
}

}  // namespace rxm
#endif
