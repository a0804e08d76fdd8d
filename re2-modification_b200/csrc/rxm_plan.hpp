// Host-side planner: decides which kernel family runs a table and builds the
// derived device tables.  No CUDA in this header.
#ifndef RXM_PLAN_HPP
#define RXM_PLAN_HPP

#include <cstdint>
#include <string>
#include <vector>

#include "rxm_host_tables.hpp"

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

constexpr uint32_t kMaxDfaStates = 4096;
constexpr int kMaxEpsDepth = 256;  // deeper == epsilon cycle: the reference overflows its stack

// RXM_OK, or RXM_ERR_UNSUPPORTED (too many DFA states / epsilon cycle) with *err set.
int plan_dfa(const rxm_tables &t, DfaPlan &out, std::string *err);

// Static checks for an MFA table (epsilon cycles, sizes).  RXM_OK or RXM_ERR_UNSUPPORTED.
int check_mfa(const rxm_tables &t, std::string *err);

}  // namespace rxm
#endif
