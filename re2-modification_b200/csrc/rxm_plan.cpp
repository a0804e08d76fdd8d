// Host-side planner (see rxm_plan.hpp).
#include "rxm_plan.hpp"

#include <cstdlib>
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
                if (!k1_classed_fits(out.n_classes, uint32_t(sets.size()) + 1u)) {
                    if (err) *err = "determinised automaton exceeds " + std::to_string(sets.size()) + " sets: its table does not fit shared memory";
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

namespace {
struct ProgGen {
    const rxm_tables &t;
    std::vector<ProgItem> &items;
    size_t root_start = 0;
    bool too_big = false;
    std::vector<std::pair<uint32_t, uint32_t>> found;  // (node, mask) pairs that can become roots

    static bool is_cell_edge(const rxm_tables &t, uint32_t e, uint32_t &k) {
        if (t.edge_kind[e] != RXM_EDGE_LIT) return false;
        const uint8_t s = t.edge_sym[e];
        if (s < '1' || s > '9') return false;
        k = uint32_t(s - '1');
        return true;
    }

    void gen(uint32_t v, uint32_t X, uint32_t C, uint32_t O, uint32_t prior, bool below_finish, int depth) {
        if (too_big) return;
        if (depth > 64 || items.size() - root_start >= kProgMaxPerRoot || items.size() >= kProgMaxItems) {
            too_big = true;
            return;
        }
        const size_t me = items.size();
        ProgItem en{};
        en.a = 0u | (below_finish ? 2u : 0u) | (v << 16);
        en.b = (C << 18);
        en.c = O | (prior << 9);
        items.push_back(en);
        found.emplace_back(v, X);
        const bool below = below_finish || v == t.finish;
        uint32_t reads_here = 0;
        bool has_leaf = false;
        for (uint32_t e = t.edge_begin[v]; e < t.edge_begin[v + 1]; e++) {
            uint32_t k = 0;
            const bool cell = is_cell_edge(t, e, k);
            if (t.edge_kind[e] == RXM_EDGE_EPS) {
                gen(t.edge_to[e], X, C, O, prior | reads_here, below, depth + 1);
            } else if (cell && !((X >> k) & 1u)) {
                const uint32_t bit = 1u << k;
                const uint32_t op = ((t.edge_open[e] >> k) & 1u) ? bit : 0u;
                gen(t.edge_to[e], X | bit, C | bit, O | op, prior | reads_here, below, depth + 1);
            } else {
                has_leaf = true;
                ProgItem lf{};
                lf.a = 1u | (below ? 2u : 0u) | (uint32_t(t.edge_kind[e]) << 4) | (uint32_t(t.edge_sym[e]) << 6) |
                       (uint32_t(t.edge_to[e]) << 16);
                lf.b = uint32_t(t.edge_open[e]) | (uint32_t(t.edge_close[e]) << 9) | (C << 18);
                lf.c = O | ((prior | reads_here) << 9) | (cell ? ((k + 1u) << 18) : 0u);
                items.push_back(lf);
                // successor's cells: everything present here plus cells its open actions create
                found.emplace_back(t.edge_to[e], X | uint32_t(t.edge_open[e]));
                if (cell) reads_here |= 1u << k;
            }
            if (too_big) return;
        }
        if (has_leaf) items[me].a |= 4u;
    }
};
}  // namespace

int compile_programs(const rxm_tables &t, MfaProgram &out, std::string *err) {
    out = MfaProgram();
    if (t.n_cells > kProgMaxCells) {
        if (err) *err = "more than " + std::to_string(kProgMaxCells) + " memory cells";
        return RXM_ERR_UNSUPPORTED;
    }
    out.n_cells = t.n_cells;
    const uint32_t nm = 1u << t.n_cells;
    out.begin.assign(size_t(t.n_states) * nm, 0xffffffffu);
    out.count.assign(size_t(t.n_states) * nm, 0);
    ProgGen g{t, out.items};
    std::vector<std::pair<uint32_t, uint32_t>> work{{t.start, 0u}};
    while (!work.empty()) {
        const auto [v, X] = work.back();
        work.pop_back();
        const size_t key = (size_t(v) << t.n_cells) | X;
        if (out.begin[key] != 0xffffffffu) continue;
        g.root_start = out.items.size();
        g.found.clear();
        out.begin[key] = uint32_t(out.items.size());
        g.gen(v, X, 0, 0, 0, false, 0);
        if (g.too_big) {
            if (err) *err = "edge program too large";
            return RXM_ERR_UNSUPPORTED;
        }
        out.count[key] = uint32_t(out.items.size() - g.root_start);
        out.max_count = std::max(out.max_count, out.count[key]);
        bool stable = (out.items[g.root_start].a & 4u) != 0;  // the root call has a leaf: re-inserted while waiting
        for (size_t x = g.root_start + 1; x < out.items.size() && stable; x++)
            if (!(out.items[x].a & 1u) && (out.items[x].a & 4u)) stable = false;  // a deeper call with a leaf
        if (stable) out.count[key] |= kProgStable;
        {   // K4's item lists: leaves, then the enters that can insert
            if (out.lbeg.empty()) {
                out.lbeg.assign(out.begin.size(), 0);
                out.lcnt.assign(out.begin.size(), 0);
            }
            out.lbeg[key] = uint32_t(out.sel.size());
            uint32_t leaves = 0, enters = 0;
            for (size_t x = g.root_start; x < out.items.size(); x++)
                if (out.items[x].a & 1u) {
                    out.sel.push_back(uint16_t(x - g.root_start));
                    leaves++;
                }
            for (size_t x = g.root_start; x < out.items.size(); x++)
                if (!(out.items[x].a & 1u) && ((out.items[x].a & 4u) || (out.items[x].a >> 16) == t.finish)) {
                    out.sel.push_back(uint16_t(x - g.root_start));
                    enters++;
                }
            out.lcnt[key] = leaves | (enters << 16);
        }
        for (const auto &f : g.found) {
            const size_t k2 = (size_t(f.first) << t.n_cells) | (f.second & (nm - 1));
            if (out.begin[k2] == 0xffffffffu) work.push_back(f);
        }
    }
    {   // K4's leaf lists by letter class: class 0 = every byte without a class of its own (a byte no literal edge
        // carries -- or, past kProgMaxClasses - 1 distinct literals, the literals that came too late: their edges
        // stay in class 0's list and are told apart by the kernel's own test of the letter)
        uint32_t nc = 1;
        for (const ProgItem &it : out.items)
            if ((it.a & 1u) && ((it.a >> 4) & 3u) == RXM_EDGE_LIT) {
                const uint32_t sym = (it.a >> 6) & 0xffu;
                if (!out.byte_class[sym] && nc < kProgMaxClasses) out.byte_class[sym] = uint8_t(nc++);
            }
        out.n_classes = nc;
        out.cbeg.assign(out.begin.size() * nc, 0);
        out.ccnt.assign(out.begin.size() * nc, 0);
        for (size_t key = 0; key < out.begin.size(); key++) {
            if (out.begin[key] == 0xffffffffu) continue;
            const uint32_t first = out.lbeg[key], leaves = out.lcnt[key] & 0xffffu;
            for (uint32_t c = 0; c < nc; c++) {
                out.cbeg[key * nc + c] = uint32_t(out.sel.size());
                uint32_t cnt = 0;
                for (uint32_t q = 0; q < leaves; q++) {
                    const uint16_t x = out.sel[first + q];
                    const ProgItem &it = out.items[out.begin[key] + x];
                    const uint32_t kind = (it.a >> 4) & 3u, sym = (it.a >> 6) & 0xffu;
                    const bool reads = ((it.c >> 18) & 0xfu) != 0;
                    if (kind == RXM_EDGE_ANY || reads || (kind == RXM_EDGE_LIT && out.byte_class[sym] == c)) {
                        out.sel.push_back(x);
                        cnt++;
                    }
                }
                out.ccnt[key * nc + c] = cnt;
            }
        }
    }
    return RXM_OK;
}

int check_nfa_bitset(const rxm_tables &t, std::string *err) {
    if (t.n_states > kBitsetMaxStates) {
        if (err) *err = "memory-free automaton with more than " + std::to_string(kBitsetMaxStates) + " states";
        return RXM_ERR_UNSUPPORTED;
    }
    // longest chain of epsilon edges (Automata::evaluateState recursion depth, automata.cpp:108-110)
    std::vector<int> depth(t.n_states, -1);  // -1 unknown, -2 in progress
    bool cyc = false;
    std::function<int(uint32_t)> dfs = [&](uint32_t q) -> int {
        if (depth[q] == -2) {
            cyc = true;
            return 0;
        }
        if (depth[q] >= 0) return depth[q];
        depth[q] = -2;
        int d = 0;
        for (uint32_t e = t.edge_begin[q]; e < t.edge_begin[q + 1] && !cyc; e++)
            if (t.edge_kind[e] == RXM_EDGE_EPS) d = std::max(d, 1 + dfs(t.edge_to[e]));
        depth[q] = d;
        return d;
    };
    int longest = 0;
    for (uint32_t q = 0; q < t.n_states && !cyc; q++) longest = std::max(longest, dfs(q));
    if (cyc) {
        if (err) *err = "epsilon cycle: the reference's evaluateState recurses without bound";
        return RXM_ERR_UNSUPPORTED;
    }
    if (uint32_t(longest) + 1 > kBitsetMaxDepth) {
        if (err) *err = "epsilon chain longer than " + std::to_string(kBitsetMaxDepth) + " edges";
        return RXM_ERR_UNSUPPORTED;
    }
    return RXM_OK;
}

void plan_bitset_masks(const rxm_tables &t, BitsetMasks &out) {
    out = BitsetMasks();
    if (t.n_states > kBitsetMaxStates) return;
    std::vector<uint8_t> letter_target(t.n_states, 0), eps_target(t.n_states, 0);
    for (uint32_t q = 0; q < t.n_states; q++)
        for (uint32_t e = t.edge_begin[q]; e < t.edge_begin[q + 1]; e++) {
            if (t.edge_kind[e] == RXM_EDGE_EPS) eps_target[t.edge_to[e]] = 1;
            else if (t.edge_kind[e] == RXM_EDGE_LIT || t.edge_kind[e] == RXM_EDGE_ANY) letter_target[t.edge_to[e]] = 1;
        }
    for (uint32_t q = 0; q < t.n_states; q++)
        if (letter_target[q] && eps_target[q]) return;  // such a node can be "visited" without being a root
    // byte classes: one per distinct literal byte, class 0 for the rest
    std::vector<int> rep{-1};
    int cls_of[256];
    std::fill(cls_of, cls_of + 256, 0);
    for (uint32_t e = 0; e < t.n_edges; e++)
        if (t.edge_kind[e] == RXM_EDGE_LIT && cls_of[t.edge_sym[e]] == 0) {
            cls_of[t.edge_sym[e]] = int(rep.size());
            rep.push_back(t.edge_sym[e]);
        }
    out.n_classes = uint32_t(rep.size());
    for (int b = 0; b < 256; b++) out.byte_class[b] = uint8_t(cls_of[b]);
    // epsilon closures (no epsilon cycle: check_nfa_bitset ran first)
    std::vector<std::vector<uint32_t>> closure(t.n_states);
    for (uint32_t r = 0; r < t.n_states; r++) {
        std::vector<uint8_t> seen(t.n_states, 0);
        std::vector<uint32_t> stack{r};
        seen[r] = 1;
        while (!stack.empty()) {
            const uint32_t u = stack.back();
            stack.pop_back();
            closure[r].push_back(u);
            for (uint32_t e = t.edge_begin[u]; e < t.edge_begin[u + 1]; e++)
                if (t.edge_kind[e] == RXM_EDGE_EPS && !seen[t.edge_to[e]]) {
                    seen[t.edge_to[e]] = 1;
                    stack.push_back(t.edge_to[e]);
                }
        }
    }
    out.ls.assign(size_t(out.n_classes) * t.n_states * 2, 0);
    for (uint32_t r = 0; r < t.n_states; r++)
        for (uint32_t u : closure[r]) {
            if (u == t.finish) out.accept[r >> 6] |= 1ull << (r & 63);
            for (uint32_t e = t.edge_begin[u]; e < t.edge_begin[u + 1]; e++) {
                const uint32_t to = t.edge_to[e];
                for (uint32_t c = 0; c < out.n_classes; c++) {
                    const bool fires = t.edge_kind[e] == RXM_EDGE_ANY ||
                                       (t.edge_kind[e] == RXM_EDGE_LIT && c != 0 && int(t.edge_sym[e]) == rep[c]);
                    if (fires) out.ls[(size_t(c) * t.n_states + r) * 2 + (to >> 6)] |= 1ull << (to & 63);
                }
            }
        }
    out.ok = true;
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
