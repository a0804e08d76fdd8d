// Owning C++ holder for an `rxm_tables` view (include/rxm.h).  Host-only.
#ifndef RXM_HOST_TABLES_HPP
#define RXM_HOST_TABLES_HPP

#include <cstdint>
#include <vector>

#include "../../include/rxm.h"

namespace rxm {

struct HostTables {
    uint32_t kind = RXM_KIND_NFA;
    uint32_t reversed = 0;
    uint32_t start = 0, finish = 0;
    uint32_t n_cells = 0;
    std::vector<uint32_t> edge_begin;  // n_states + 1
    std::vector<uint8_t> edge_kind, edge_sym;
    std::vector<uint16_t> edge_to, edge_open, edge_close;

    uint32_t n_states() const { return edge_begin.empty() ? 0 : uint32_t(edge_begin.size() - 1); }
    uint32_t n_edges() const { return uint32_t(edge_kind.size()); }

    // The returned view borrows this object's storage.
    rxm_tables view() const {
        rxm_tables t{};
        t.abi_version = RXM_ABI_VERSION;
        t.kind = kind;
        t.reversed = reversed;
        t.n_states = n_states();
        t.n_edges = n_edges();
        t.start = start;
        t.finish = finish;
        t.n_cells = n_cells;
        t.edge_begin = edge_begin.data();
        t.edge_kind = edge_kind.data();
        t.edge_sym = edge_sym.data();
        t.edge_to = edge_to.data();
        t.edge_open = edge_open.data();
        t.edge_close = edge_close.data();
        return t;
    }

    void assign(const rxm_tables &t) {
        kind = t.kind;
        reversed = t.reversed;
        start = t.start;
        finish = t.finish;
        n_cells = t.n_cells;
        edge_begin.assign(t.edge_begin, t.edge_begin + t.n_states + 1);
        edge_kind.assign(t.edge_kind, t.edge_kind + t.n_edges);
        edge_sym.assign(t.edge_sym, t.edge_sym + t.n_edges);
        edge_to.assign(t.edge_to, t.edge_to + t.n_edges);
        edge_open.assign(t.edge_open, t.edge_open + t.n_edges);
        edge_close.assign(t.edge_close, t.edge_close + t.n_edges);
    }
};

}  // namespace rxm
#endif
