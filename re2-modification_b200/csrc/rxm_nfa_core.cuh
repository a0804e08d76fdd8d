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

// The same step from follow masks (rxm_plan.hpp: BitsetMasks): ls = LS[class] = n_states x 2 words (16 bytes per
// state, 16-byte aligned).  The sets are walked as four 32-bit words: a root in word W filters the words below W with
// all of S, its own word with the members below it, and the words above not at all -- per member one find-first,
// one 16-byte load and four three-input logic operations (the 64-bit form cost ~25 instructions per member).
struct NfaRow {
    uint32_t x, y, z, w;
};
RXM_HD NfaRow nfa_mask_row(const uint64_t *ls, uint32_t r) {
#if defined(__CUDA_ARCH__)
    const uint4 v = *reinterpret_cast<const uint4 *>(ls + 2u * r);
    return NfaRow{v.x, v.y, v.z, v.w};
#else
    const uint64_t lo = ls[2u * r], hi = ls[2u * r + 1u];
    return NfaRow{uint32_t(lo), uint32_t(lo >> 32), uint32_t(hi), uint32_t(hi >> 32)};
#endif
}
RXM_HD uint32_t nfa_ctz32(uint32_t w) {
#if defined(__CUDA_ARCH__)
    return uint32_t(__ffs(int(w)) - 1);
#else
    return uint32_t(__builtin_ctz(w));
#endif
}
RXM_HD Bits128 nfa_mask_step(const uint64_t *ls, Bits128 S) {
    const uint32_t s0 = uint32_t(S.lo), s1 = uint32_t(S.lo >> 32), s2 = uint32_t(S.hi), s3 = uint32_t(S.hi >> 32);
    uint32_t n0 = 0, n1 = 0, n2 = 0, n3 = 0;
    for (uint32_t w = s0; w; w &= w - 1u) {
        const uint32_t r = nfa_ctz32(w);
        const NfaRow L = nfa_mask_row(ls, r);
        n0 |= L.x & ~(s0 & ((1u << r) - 1u));
        n1 |= L.y;
        n2 |= L.z;
        n3 |= L.w;
    }
    for (uint32_t w = s1; w; w &= w - 1u) {
        const uint32_t r = nfa_ctz32(w);
        const NfaRow L = nfa_mask_row(ls, r + 32u);
        n0 |= L.x & ~s0;
        n1 |= L.y & ~(s1 & ((1u << r) - 1u));
        n2 |= L.z;
        n3 |= L.w;
    }
    for (uint32_t w = s2; w; w &= w - 1u) {
        const uint32_t r = nfa_ctz32(w);
        const NfaRow L = nfa_mask_row(ls, r + 64u);
        n0 |= L.x & ~s0;
        n1 |= L.y & ~s1;
        n2 |= L.z & ~(s2 & ((1u << r) - 1u));
        n3 |= L.w;
    }
    for (uint32_t w = s3; w; w &= w - 1u) {
        const uint32_t r = nfa_ctz32(w);
        const NfaRow L = nfa_mask_row(ls, r + 96u);
        n0 |= L.x & ~s0;
        n1 |= L.y & ~s1;
        n2 |= L.z & ~s2;
        n3 |= L.w & ~(s3 & ((1u << r) - 1u));
    }
    return Bits128{uint64_t(n0) | (uint64_t(n1) << 32), uint64_t(n2) | (uint64_t(n3) << 32)};
}

}  // namespace rxm
#endif
