// Host-side planner (see rxm_plan.hpp).
#include "rxm_plan.hpp"

#include <algorithm>
#include <functional>
#include <map>

namespace rxm {

namespace {

using StateSet = std::vector<uint16_t>;  // sorted ascending == std::set<Node*> order

// One Automata::evaluateStates call (automata.cpp:119-128) with evaluateState
// (automata.cpp:98-117) on the flat table.  `letter` < 0 is the final pass with
// letter "" (automata.cpp:201-202).  Returns false on runaway recursion.
struct ExactStep {
    const rxm_tables &t;
    std::vector<uint8_t> visited, next;
    int letter = 0;
    bool ok = true;

    explicit ExactStep(const rxm_tables &tt) : t(tt), visited(tt.n_states), next(tt.n_states) {}

    void eval(uint32_t q, int depth) {
        if (!ok) return;
        if (depth > kMaxEpsDepth) {
            ok = false;
            return;
        }
        if (letter < 0 && q == t.finish) {
            next[q] = 1;
        } else {
            for (uint32_t e = t.edge_begin[q]; e < t.edge_begin[q + 1]; e++) {
                const uint32_t to = t.edge_to[e];
                if (visited[to]) continue;
                const uint8_t kind = t.edge_kind[e];
                if (kind == RXM_EDGE_EPS) eval(to, depth + 1);
                else if (letter >= 0 && (kind == RXM_EDGE_ANY ||
                                         (kind == RXM_EDGE_LIT && t.edge_sym[e] == uint8_t(letter))))
                    next[to] = 1;
            }
        }
        visited[q] = 1;
    }

    bool run(const StateSet &s, int ch, StateSet &out) {
        letter = ch;
        std::fill(visited.begin(), visited.end(), 0);
        std::fill(next.begin(), next.end(), 0);
        for (uint16_t q : s)
            if (!visited[q]) eval(q, 0);
        out.clear();
        for (uint32_t q = 0; q < t.n_states; q++)
            if (next[q]) out.push_back(uint16_t(q));
        return ok;
    }
};

// Textbook step on the same representation (letter-move targets, not yet closed).
struct CleanStep {
    const rxm_tables &t;
    std::vector<uint8_t> closed;
    explicit CleanStep(const rxm_tables &tt) : t(tt), closed(tt.n_states) {}

    void closure(const StateSet &s) {
        std::fill(closed.begin(), closed.end(), 0);
        std::vector<uint16_t> stack(s.begin(), s.end());
        for (uint16_t q : s) closed[q] = 1;
        while (!stack.empty()) {
            uint16_t q = stack.back();
            stack.pop_back();
            for (uint32_t e = t.edge_begin[q]; e < t.edge_begin[q + 1]; e++)
                if (t.edge_kind[e] == RXM_EDGE_EPS && !closed[t.edge_to[e]]) {
                    closed[t.edge_to[e]] = 1;
                    stack.push_back(t.edge_to[e]);
                }
        }
    }
    void run(const StateSet &s, int ch, StateSet &out) {
        closure(s);
        std::vector<uint8_t> next(t.n_states, 0);
        if (ch < 0) {
            next[t.finish] = closed[t.finish];
        } else {
            for (uint32_t q = 0; q < t.n_states; q++) {
                if (!closed[q]) continue;
                for (uint32_t e = t.edge_begin[q]; e < t.edge_begin[q + 1]; e++) {
                    const uint8_t kind = t.edge_kind[e];
                    if (kind == RXM_EDGE_ANY || (kind == RXM_EDGE_LIT && t.edge_sym[e] == uint8_t(ch)))
                        next[t.edge_to[e]] = 1;
                }
            }
        }
        out.clear();
        for (uint32_t q = 0; q < t.n_states; q++)
            if (next[q]) out.push_back(uint16_t(q));
    }
};

}  // namespace

int plan_dfa(const rxm_tables &t, DfaPlan &out, std::string *err) {
    out = DfaPlan();
    out.reversed = t.reversed ? 1 : 0;

    // byte classes: one per distinct literal byte, class 0 for all other bytes
    int cls_of[256];
    std::fill(cls_of, cls_of + 256, 0);
    std::vector<int> rep{-2};  // representative byte per class; class 0 fixed below
    for (uint32_t e = 0; e < t.n_edges; e++)
        if (t.edge_kind[e] == RXM_EDGE_LIT && cls_of[t.edge_sym[e]] == 0) {
            cls_of[t.edge_sym[e]] = int(rep.size());
            rep.push_back(t.edge_sym[e]);
        }
    for (int b = 0; b < 256; b++)
        if (cls_of[b] == 0) {
            rep[0] = b;
            break;
        }
    // (if all 256 bytes are literals, class 0 is empty; rep[0] stays -2 and is never used)
    out.n_classes = uint32_t(rep.size());
    for (int b = 0; b < 256; b++) out.byte_class[b] = uint8_t(cls_of[b]);
    if (out.n_classes > 256) return RXM_ERR_UNSUPPORTED;

    ExactStep exact(t);
    CleanStep clean(t);
    std::map<StateSet, uint32_t> ids;
    std::vector<StateSet> sets;
    sets.push_back(StateSet());  // 0 = dead
    ids[sets[0]] = 0;
    StateSet s0{uint16_t(t.start)};
    ids[s0] = 1;
    sets.push_back(s0);
    out.start = 1;
    std::vector<std::vector<uint16_t>> rows;  // rows[state][class]
    StateSet nx, nc;
    for (uint32_t cur = 0; cur < sets.size(); cur++) {
        std::vector<uint16_t> row(out.n_classes, 0);
        const StateSet s = sets[cur];
        for (uint32_t c = 0; c < out.n_classes; c++) {
            if (cur == 0 || rep[c] < 0) {
                row[c] = 0;
                continue;
            }
            if (!exact.run(s, rep[c], nx)) {
                if (err) *err = "epsilon cycle: the reference's evaluateState recurses without bound";
                return RXM_ERR_UNSUPPORTED;
            }
            clean.run(s, rep[c], nc);
            if (nx != nc) out.exact_step_differs++;
            auto it = ids.find(nx);
            if (it == ids.end()) {
                if (sets.size() >= kMaxDfaStates) {
                    if (err) *err = "determinised automaton exceeds " + std::to_string(kMaxDfaStates) + " states";
                    return RXM_ERR_UNSUPPORTED;
                }
                it = ids.emplace(nx, uint32_t(sets.size())).first;
                sets.push_back(nx);
            }
            row[c] = uint16_t(it->second);
        }
        rows.push_back(row);
        // acceptance: final pass with letter ""
        uint8_t acc = 0;
        if (cur != 0) {
            if (!exact.run(s, -1, nx)) {
                if (err) *err = "epsilon cycle: the reference's evaluateState recurses without bound";
                return RXM_ERR_UNSUPPORTED;
            }
            acc = std::binary_search(nx.begin(), nx.end(), uint16_t(t.finish)) ? 1 : 0;
            clean.run(s, -1, nc);
            const bool cacc = std::binary_search(nc.begin(), nc.end(), uint16_t(t.finish));
            if (cacc != bool(acc)) out.exact_step_differs++;
        }
        out.accept.push_back(acc);
    }
    out.n_states = uint32_t(sets.size());
    out.trans.assign(size_t(out.n_classes) * out.n_states, 0);
    for (uint32_t s = 0; s < out.n_states; s++)
        for (uint32_t c = 0; c < out.n_classes; c++) out.trans[size_t(c) * out.n_states + s] = rows[s][c];
    return RXM_OK;
}

int check_mfa(const rxm_tables &t, std::string *err) {
    // epsilon-only cycles make MFA::evaluateState (mfa.cpp:143-147) recurse forever
    std::vector<uint8_t> color(t.n_states, 0);
    bool cyc = false;
    std::function<void(uint32_t)> dfs = [&](uint32_t q) {
        color[q] = 1;
        for (uint32_t e = t.edge_begin[q]; e < t.edge_begin[q + 1] && !cyc; e++) {
            if (t.edge_kind[e] != RXM_EDGE_EPS) continue;
            const uint32_t to = t.edge_to[e];
            if (color[to] == 1) cyc = true;
            else if (color[to] == 0) dfs(to);
        }
        color[q] = 2;
    };
    for (uint32_t q = 0; q < t.n_states && !cyc; q++)
        if (!color[q]) dfs(q);
    if (cyc) {
        if (err) *err = "epsilon cycle: the reference's MFA::evaluateState recurses without bound";
        return RXM_ERR_UNSUPPORTED;
    }
    return RXM_OK;
}

}  // namespace rxm
