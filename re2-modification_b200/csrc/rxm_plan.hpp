    // The same leaves split by the input letter's CLASS (class 0: the bytes without a class of their own; classes
    // 1 .. n_classes-1: distinct literal bytes, at most kProgMaxClasses - 1 of them): cbeg / ccnt[key * n_classes + c]
    // delimit, in `sel`, the leaves that can act on a letter of class c -- `.` edges, literal edges of that class,
    // reads of a present cell -- in program order (a subsequence of the list above, so ties are decided as
    // before).  A literal edge of another byte never fires on that letter: about half of an active
    // configuration's items are not looked at.  n_classes >= 1.
// Host-side planner: decides which kernel family runs a table and builds the
// derived device tables.  No CUDA in this header.
#ifndef RXM_PLAN_HPP
#define RXM_PLAN_HPP

#include <cstdint>
#include <string>
#include <vector>

#include "rxm_host_tables.hpp"
#include "rxm_mfa_core.cuh"  // ProgItem

namespace rxm {

// Determinised memory-free automaton ("K1").  DFA states are the active sets
// (std::set<Node*> contents, automata.cpp:178-200) reachable from {start} under
// the reference's EXACT step -- including its `visited` filtering of letter
// edges (automata.cpp:104-107) -- so the DFA reproduces the reference's answers
// even where they differ from the textbook NFA.  State 0 is the empty set
// (absorbing: automata.cpp:186-188 breaks out of the loop there).
struct DfaPlan {
    uint32_t n_states = 0;    // including the dead state 0
    uint32_t n_classes = 0;   // byte classes; class 0 = "every byte no edge names"
    uint32_t start = 0;
    uint32_t reversed = 0;
    uint32_t exact_step_differs = 0;
    uint8_t byte_class[256] = {0};
    std::vector<uint16_t> trans;  // [n_classes][n_states]
    std::vector<uint8_t> accept;  // [n_states]: finish in the set after the final epsilon pass (:201-209)
};

// The determinisation goes on for as long as its table still fits K1's shared memory in the two-lookup form
// (byte class u8 [256], then [class][set] u16, then one accept byte per set rounded up to 256): ~25 000 sets with
// three byte classes.  (Round 1 stopped at 4096 sets; the 24 577 sets of `(a|b)*a(a|b)^12 b(a|b)*` went to the
// bit-set engine at 38 GB/s although their table is 172 KB.)  Set ids are u16: 65 535 at most.
constexpr uint32_t kMaxDfaStates = 65535;
// table + accept bytes K1's two-lookup kernel can keep next to its ring: 216 KB of dynamic shared memory less the
// ring (8 warps x 2 stages x 32 lanes x 80 bytes) and the ring's alignment (rxm_k1.cu: launch_classed asserts it)
constexpr size_t kK1ClassedBytes = 216 * 1024 - 8 * 2 * 32 * 80 - 128;
constexpr size_t k1_classed_table_bytes(uint32_t n_classes, uint32_t n_states) {
    return 256 + 2 * size_t(n_classes) * n_states;
}
constexpr size_t k1_accept_bytes(uint32_t n_states) { return (size_t(n_states) + 255) & ~size_t(255); }
constexpr bool k1_classed_fits(uint32_t n_classes, uint32_t n_states) {
    return n_states <= kMaxDfaStates && k1_classed_table_bytes(n_classes, n_states) + k1_accept_bytes(n_states) <= kK1ClassedBytes;
}
constexpr int kMaxEpsDepth = 256;  // deeper == epsilon cycle: the reference overflows its stack

// RXM_OK, or RXM_ERR_UNSUPPORTED (too many DFA states / epsilon cycle) with *err set.
int plan_dfa(const rxm_tables &t, DfaPlan &out, std::string *err);

// ---- MFA "edge programs" ---------------------------------------------------------------
// MFA::evaluateState (mfa.cpp:136-200) recurses through epsilon edges and through read
// edges whose cell is absent (mfa.cpp:143-160).  Which of those recursions happen depends
// only on WHICH cells the configuration has, so for every reachable (node, set of existing
// cells) the whole depth-first walk is flattened on the host into a list of items in the
// reference's visiting order:
//   ENTER  a call of evaluateState on node v (the root call or a recursive one)
//   LEAF   a non-recursive edge met in that call (letter, '.', never, read of a present cell)
// with, per item, the cells created on the way down (absent-cell edges), which cells earlier
// read edges of the enclosing calls have marked is_read (Variable::read, mfa.cpp:177), and
// whether the item lies below a call on `finish` (skipped when first == len, mfa.cpp:138).
// The item index doubles as the creation order of everything the walk allocates.
constexpr uint32_t kProgMaxCells = 4;      // programs are keyed by (node, 2^cells masks)
constexpr uint32_t kProgMaxItems = 1u << 16;
constexpr uint32_t kProgMaxPerRoot = 4096;

struct MfaProgram {
    uint32_t n_cells = 0;
    std::vector<ProgItem> items;
    std::vector<uint32_t> begin;  // [node << n_cells | mask] -> first item, 0xffffffff if unreachable
    std::vector<uint32_t> count;  // [node << n_cells | mask] -> number of items
    uint32_t max_count = 0;
    // K4 (thread per string) walks only the items that can act for the configuration in hand: per
    // key, sel[lbeg .. lbeg + leaves) are the LEAF items (an ACTIVE configuration, mfa.cpp:161-193)
    // and the next `enters` entries the ENTER items that can insert (a WAITING / final one,
    // mfa.cpp:138-140, 195-197: the call has a leaf, or it is a call on `finish`); entries are item
    // indices relative to begin[key], in program order.  lcnt[key] = leaves | enters << 16.
    std::vector<uint32_t> lbeg, lcnt;
    std::vector<uint16_t> sel;
    // The same leaves split by the input letter's CLASS (class 0: a byte no literal edge carries; classes
    // 1 .. n_classes-1: the distinct literal bytes): cbeg / ccnt[key * n_classes + c] delimit, in `sel`, the
    // leaves that can act on a letter of class c -- `.` edges, literal edges of that byte, reads of a present
    // cell -- in program order (a subsequence of the list above, so ties are decided as before).  A literal
    // edge of another byte never fires on that letter: about half of an active configuration's items are not
    // looked at.  n_classes == 0: too many distinct literals, K4 walks the undivided list.
    uint32_t n_classes = 0;
    std::vector<uint32_t> cbeg, ccnt;
    uint8_t byte_class[256] = {0};
};
constexpr uint32_t kProgMaxClasses = 8;
// K4Prog::lists (rxm_k4_core.cuh): begin | count | lbeg | lcnt | cbeg | ccnt | byte_class, one array
inline std::vector<uint32_t> k4_pack_lists(const MfaProgram &p) {
    const size_t nk = p.begin.size();
    std::vector<uint32_t> v((4 + 2 * size_t(p.n_classes)) * nk + 64, 0);
    for (size_t k = 0; k < nk; k++) {
        v[k] = p.begin[k];
        v[nk + k] = p.count[k];
        v[2 * nk + k] = p.lbeg.empty() ? 0u : p.lbeg[k];
        v[3 * nk + k] = p.lcnt.empty() ? 0u : p.lcnt[k];
    }
    for (size_t k = 0; k < p.cbeg.size(); k++) {
        v[4 * nk + k] = p.cbeg[k];
        v[(4 + p.n_classes) * nk + k] = p.ccnt[k];
    }
    uint8_t *cls = reinterpret_cast<uint8_t *>(v.data() + (4 + 2 * size_t(p.n_classes)) * nk);
    for (int b = 0; b < 256; b++) cls[b] = p.byte_class[b];
    return v;
}

// RXM_OK, or RXM_ERR_UNSUPPORTED (more than kProgMaxCells cells, program too large).
int compile_programs(const rxm_tables &t, MfaProgram &out, std::string *err);

// Memory-free automaton simulated as a bit set ("K1_BITSET", rxm_k1b.cu): used when the exact-
// step determinisation above exceeds kMaxDfaStates.  Limits: kBitsetMaxStates states, no
// epsilon cycle (the reference overflows its stack there), epsilon chains of at most
// kBitsetMaxDepth edges (the kernel's explicit recursion stack).
constexpr uint32_t kBitsetMaxStates = 128;
constexpr uint32_t kBitsetMaxDepth = 48;
int check_nfa_bitset(const rxm_tables &t, std::string *err);

// Follow masks for K1B's bit-parallel step.  When no letter-edge target has an incoming epsilon
// edge (true for everything toGlushkov and toThomson build: a letter edge always leads to a fresh
// node) the reference's step collapses to
//     next = OR over r in S of  LS[class(letter)][r] & ~(S & bits_below(r))
// with LS[c][r] = letter-c successors of r's epsilon closure: a node is only ever "visited" as a
// completed ROOT of the walk (roots run in ascending order, automata.cpp:122), so the `visited`
// test of automata.cpp:105-107 removes exactly the targets that sit in S below the current root,
// and a node reached from two roots is evaluated under the smaller one, whose filter is the
// weaker.  Acceptance (final pass, :100-102, :201-209): finish lies in the closure of some r in S.
struct BitsetMasks {
    bool ok = false;               // the structural condition holds; otherwise K1B walks the edges
    uint32_t n_classes = 0;        // class 0: bytes no literal edge names (only `.` edges fire)
    uint8_t byte_class[256] = {0};
    std::vector<uint64_t> ls;      // [n_classes][n_states][2]
    uint64_t accept[2] = {0, 0};   // r: finish in eclose(r)
};
void plan_bitset_masks(const rxm_tables &t, BitsetMasks &out);

// Static checks for an MFA table (epsilon cycles, sizes).  RXM_OK or RXM_ERR_UNSUPPORTED.
int check_mfa(const rxm_tables &t, std::string *err);

}  // namespace rxm
#endif
