// Flattening stage: reference object graph -> rxm_tables.
//
// Runs IN the process that called Regexp::compile (regex/regex.cpp:266-343) and
// reads the automaton it returned:
//   Automata{start, finish, is_reversed}        automata.h:18-48
//   MFA{start, finish, is_reversed} (shadowing) automata.h:50-84
//   Node{edges} / MemoryNode{edges}             node.h:11-38
//   Edge{by,to} / MemoryEdge{by,to,memoryActions}  edge.h:13-49
//
// This header is the only product file that includes reference headers; it is
// compiled with -I<reference root> (see re2-modification_b200/Makefile).  It
// never calls MFA::match / Automata::match.
//
// Numbering: states are the nodes reachable from `start` (the `nodes` lists are
// merged/erased ad hoc by the builders, bt/bt_mfa.cpp:50-51,90-93, so they are
// not trusted) plus `finish`, numbered by ADDRESS RANK -- the order in which
// std::set<Node*> (automata.cpp:122) and std::set<MemoryState> (mfa.cpp:206)
// iterate them.  Edges keep std::list order.
#ifndef RXM_FLATTEN_HPP
#define RXM_FLATTEN_HPP

#include <algorithm>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "automata.h"  // reference header, via -I<reference root>

#include "rxm_host_tables.hpp"

namespace rxm {

namespace detail {

inline bool classify_label(const std::string &by, bool is_mfa, uint8_t &kind, uint8_t &sym,
                           std::string *err) {
    sym = 0;
    if (by.empty() || (is_mfa && by == "\xCE\xB5")) {  // "" or "ε" (mfa.cpp:143)
        kind = RXM_EDGE_EPS;
    } else if (by == ".") {
        kind = RXM_EDGE_ANY;
    } else if (by.size() == 1) {
        kind = RXM_EDGE_LIT;
        sym = uint8_t(by[0]);
    } else {
        // e.g. `string s(&rune)` picking up bytes after the rune (bt_thomson.cpp:11):
        // never equal to a one-byte input letter, so the edge can never fire.
        kind = RXM_EDGE_NEVER;
        if (is_mfa && std::string("1") <= by && by <= std::string("9")) {
            // mfa.cpp:148 would treat this multi-byte label as a cell name.
            if (err) *err = "multi-character memory cell name '" + by + "' is not supported";
            return false;
        }
    }
    return true;
}

template <class NodeT, class EdgesOf, class ToOf>
void collect_reachable(NodeT *start, NodeT *finish, EdgesOf edges_of, ToOf to_of,
                       std::vector<NodeT *> &out) {
    std::vector<NodeT *> stack{start};
    std::map<NodeT *, bool, std::less<NodeT *>> seen;
    seen[start] = true;
    while (!stack.empty()) {
        NodeT *q = stack.back();
        stack.pop_back();
        for (auto *e : edges_of(q)) {
            NodeT *to = to_of(e);
            if (to && !seen[to]) {
                seen[to] = true;
                stack.push_back(to);
            }
        }
    }
    if (finish) seen[finish] = true;
    out.clear();
    for (auto &kv : seen)
        if (kv.second) out.push_back(kv.first);
    std::sort(out.begin(), out.end(), std::less<NodeT *>());  // address rank
}

}  // namespace detail

// Returns RXM_OK, RXM_ERR_INVALID or RXM_ERR_UNSUPPORTED.
inline int flatten(Automata *automata, bool is_mfa, HostTables &out, std::string *err = nullptr) {
    if (!automata) return RXM_ERR_INVALID;
    out = HostTables();
    if (is_mfa) {
        MFA *mfa = static_cast<MFA *>(automata);  // match.cpp:17-19
        std::vector<MemoryNode *> nodes;
        detail::collect_reachable<MemoryNode>(
            mfa->start, mfa->finish, [](MemoryNode *q) -> std::list<MemoryEdge *> & { return q->edges; },
            [](MemoryEdge *e) { return e->to; }, nodes);
        if (nodes.size() > RXM_MAX_STATES) return RXM_ERR_UNSUPPORTED;
        std::map<MemoryNode *, uint32_t> id;
        for (uint32_t i = 0; i < nodes.size(); i++) id[nodes[i]] = i;
        out.kind = RXM_KIND_MFA;
        out.reversed = mfa->is_reversed ? 1 : 0;  // MFA::is_reversed, automata.h:55
        out.start = id[mfa->start];
        out.finish = id[mfa->finish];
        out.edge_begin.push_back(0);
        for (MemoryNode *q : nodes) {
            for (MemoryEdge *e : q->edges) {
                uint8_t kind, sym;
                if (!detail::classify_label(e->by, true, kind, sym, err)) return RXM_ERR_UNSUPPORTED;
                uint16_t open_mask = 0, close_mask = 0;
                for (const auto &act : e->memoryActions) {  // edge.h:36
                    const std::string &var = act.first;
                    if (var.size() != 1 || var[0] < '1' || var[0] > '9') {
                        if (err) *err = "memory cell name '" + var + "' outside 1..9";
                        return RXM_ERR_UNSUPPORTED;
                    }
                    const uint32_t k = uint32_t(var[0] - '1');
                    out.n_cells = std::max(out.n_cells, k + 1);
                    if (act.second == open) open_mask |= uint16_t(1u << k);
                    else close_mask |= uint16_t(1u << k);
                }
                if (kind == RXM_EDGE_LIT && sym >= '1' && sym <= '9')
                    out.n_cells = std::max(out.n_cells, uint32_t(sym - '1') + 1);
                out.edge_kind.push_back(kind);
                out.edge_sym.push_back(sym);
                out.edge_to.push_back(uint16_t(id[e->to]));
                out.edge_open.push_back(open_mask);
                out.edge_close.push_back(close_mask);
            }
            out.edge_begin.push_back(uint32_t(out.edge_kind.size()));
        }
    } else {
        std::vector<Node *> nodes;
        detail::collect_reachable<Node>(
            automata->start, automata->finish,
            [](Node *q) -> std::list<Edge *> & { return q->edges; }, [](Edge *e) { return e->to; },
            nodes);
        if (nodes.size() > RXM_MAX_STATES) return RXM_ERR_UNSUPPORTED;
        std::map<Node *, uint32_t> id;
        for (uint32_t i = 0; i < nodes.size(); i++) id[nodes[i]] = i;
        out.kind = RXM_KIND_NFA;
        out.reversed = automata->is_reversed ? 1 : 0;
        out.start = id[automata->start];
        out.finish = id[automata->finish];
        out.edge_begin.push_back(0);
        for (Node *q : nodes) {
            for (Edge *e : q->edges) {
                uint8_t kind, sym;
                if (!detail::classify_label(e->by, false, kind, sym, err)) return RXM_ERR_UNSUPPORTED;
                out.edge_kind.push_back(kind);
                out.edge_sym.push_back(sym);
                out.edge_to.push_back(uint16_t(id[e->to]));
                out.edge_open.push_back(0);
                out.edge_close.push_back(0);
            }
            out.edge_begin.push_back(uint32_t(out.edge_kind.size()));
        }
    }
    return RXM_OK;
}

}  // namespace rxm
#endif
