// Host-side table utilities of the C ABI (include/rxm.h): validate, text
// format / parse, release.  No CUDA here.
//
// Text form (one automaton):
//   rxm-tables 1
//   kind mfa|nfa
//   reversed 0|1
//   states N
//   start S
//   finish F
//   cells C
//   edges E
//   <from> <E|L|A|N> <sym> <to> [o<k>|c<k>]...      (E lines, grouped by <from>)
//   end
// <sym> is '-' (none), one character of [A-Za-z0-9], or #<decimal byte>.
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <sstream>
#include <string>
#include <vector>

#include "rxm_host_tables.hpp"

extern "C" int rxm_tables_validate(const rxm_tables *t) {
    if (!t) return RXM_ERR_INVALID;
    if (t->abi_version != RXM_ABI_VERSION) return RXM_ERR_INVALID;
    if (t->kind != RXM_KIND_NFA && t->kind != RXM_KIND_MFA) return RXM_ERR_INVALID;
    if (t->n_states == 0 || t->n_states > RXM_MAX_STATES) return RXM_ERR_INVALID;
    if (t->start >= t->n_states || t->finish >= t->n_states) return RXM_ERR_INVALID;
    if (t->n_cells > RXM_MAX_CELLS) return RXM_ERR_INVALID;
    if (!t->edge_begin) return RXM_ERR_INVALID;
    if (t->n_edges && (!t->edge_kind || !t->edge_sym || !t->edge_to || !t->edge_open || !t->edge_close))
        return RXM_ERR_INVALID;
    if (t->edge_begin[0] != 0 || t->edge_begin[t->n_states] != t->n_edges) return RXM_ERR_INVALID;
    for (uint32_t q = 0; q < t->n_states; q++)
        if (t->edge_begin[q] > t->edge_begin[q + 1]) return RXM_ERR_INVALID;
    const uint16_t cell_mask = uint16_t((1u << t->n_cells) - 1u);
    for (uint32_t e = 0; e < t->n_edges; e++) {
        if (t->edge_kind[e] > RXM_EDGE_NEVER) return RXM_ERR_INVALID;
        if (t->edge_to[e] >= t->n_states) return RXM_ERR_INVALID;
        if (t->edge_open[e] & t->edge_close[e]) return RXM_ERR_INVALID;
        if ((t->edge_open[e] | t->edge_close[e]) & ~cell_mask) return RXM_ERR_INVALID;
        if (t->kind == RXM_KIND_NFA && (t->edge_open[e] | t->edge_close[e])) return RXM_ERR_INVALID;
        if (t->kind == RXM_KIND_MFA && t->edge_kind[e] == RXM_EDGE_LIT && t->edge_sym[e] >= '1' &&
            t->edge_sym[e] <= '9' && uint32_t(t->edge_sym[e] - '0') > t->n_cells)
            return RXM_ERR_INVALID;
    }
    return RXM_OK;
}

static int tables_format_impl(const rxm_tables *t, char *buf, size_t buf_size, size_t *needed) {
    int st = rxm_tables_validate(t);
    if (st != RXM_OK) return st;
    std::ostringstream os;
    os << "rxm-tables 1\n";
    os << "kind " << (t->kind == RXM_KIND_MFA ? "mfa" : "nfa") << "\n";
    os << "reversed " << (t->reversed ? 1 : 0) << "\n";
    os << "states " << t->n_states << "\n";
    os << "start " << t->start << "\n";
    os << "finish " << t->finish << "\n";
    os << "cells " << t->n_cells << "\n";
    os << "edges " << t->n_edges << "\n";
    static const char kinds[] = {'E', 'L', 'A', 'N'};
    for (uint32_t q = 0; q < t->n_states; q++) {
        for (uint32_t e = t->edge_begin[q]; e < t->edge_begin[q + 1]; e++) {
            os << q << ' ' << kinds[t->edge_kind[e]] << ' ';
            const uint8_t s = t->edge_sym[e];
            if (t->edge_kind[e] != RXM_EDGE_LIT) os << '-';
            else if (std::isalnum(s)) os << char(s);
            else os << '#' << unsigned(s);
            os << ' ' << t->edge_to[e];
            for (uint32_t k = 0; k < RXM_MAX_CELLS; k++) {
                if ((t->edge_open[e] >> k) & 1) os << " o" << (k + 1);
                if ((t->edge_close[e] >> k) & 1) os << " c" << (k + 1);
            }
            os << "\n";
        }
    }
    os << "end\n";
    const std::string s = os.str();
    if (needed) *needed = s.size() + 1;
    if (buf && buf_size) {
        const size_t n = s.size() < buf_size - 1 ? s.size() : buf_size - 1;
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
        if (n < s.size()) return RXM_ERR_INVALID;  // truncated
    }
    return RXM_OK;
}

namespace {
struct Block {  // one allocation: header + arrays
    rxm_tables t;
};
}  // namespace

static int tables_parse_impl(const char *text, size_t len, rxm_tables **out) {
    if (!text || !out) return RXM_ERR_INVALID;
    *out = nullptr;
    std::istringstream is(std::string(text, len));
    std::string w, v;
    rxm::HostTables h;
    uint32_t states = 0, edges = 0;
    if (!(is >> w >> v) || w != "rxm-tables" || v != "1") return RXM_ERR_PARSE;
    auto kv = [&](const char *key, std::string &val) { return bool(is >> w >> val) && w == key; };
    if (!kv("kind", v)) return RXM_ERR_PARSE;
    if (v == "mfa") h.kind = RXM_KIND_MFA;
    else if (v == "nfa") h.kind = RXM_KIND_NFA;
    else return RXM_ERR_PARSE;
    auto num = [&](const char *key, uint32_t &val) {
        std::string s;
        if (!kv(key, s) || s.empty()) return false;
        char *endp = nullptr;
        unsigned long x = std::strtoul(s.c_str(), &endp, 10);
        if (*endp) return false;
        val = uint32_t(x);
        return true;
    };
    if (!num("reversed", h.reversed) || !num("states", states) || !num("start", h.start) ||
        !num("finish", h.finish) || !num("cells", h.n_cells) || !num("edges", edges))
        return RXM_ERR_PARSE;
    if (states == 0 || states > RXM_MAX_STATES || edges > (1u << 24)) return RXM_ERR_PARSE;
    std::string line;
    std::getline(is, line);  // rest of the "edges" line
    h.edge_begin.assign(states + 1, 0);
    uint32_t cur_from = 0, seen = 0;
    bool got_end = false;
    while (std::getline(is, line)) {
        if (line.empty()) continue;
        if (line == "end") {
            got_end = true;
            break;
        }
        std::istringstream ls(line);
        uint32_t from, to;
        std::string kind, sym, act;
        if (!(ls >> from >> kind >> sym >> to) || kind.size() != 1) return RXM_ERR_PARSE;
        if (from >= states || to >= states || from < cur_from) return RXM_ERR_PARSE;  // (before `to` is narrowed)
        while (cur_from < from) h.edge_begin[++cur_from] = seen;
        uint8_t k;
        switch (kind[0]) {
            case 'E': k = RXM_EDGE_EPS; break;
            case 'L': k = RXM_EDGE_LIT; break;
            case 'A': k = RXM_EDGE_ANY; break;
            case 'N': k = RXM_EDGE_NEVER; break;
            default: return RXM_ERR_PARSE;
        }
        uint8_t s = 0;
        if (k == RXM_EDGE_LIT) {
            if (sym.size() == 1 && std::isalnum((unsigned char)sym[0])) s = uint8_t(sym[0]);
            else if (sym.size() >= 2 && sym[0] == '#') {
                char *endp = nullptr;
                unsigned long x = std::strtoul(sym.c_str() + 1, &endp, 10);
                if (*endp || x > 255) return RXM_ERR_PARSE;
                s = uint8_t(x);
            } else return RXM_ERR_PARSE;
        } else if (sym != "-") return RXM_ERR_PARSE;
        uint16_t om = 0, cm = 0;
        while (ls >> act) {
            if (act.size() != 2 || act[1] < '1' || act[1] > '9') return RXM_ERR_PARSE;
            if (act[0] == 'o') om |= uint16_t(1u << (act[1] - '1'));
            else if (act[0] == 'c') cm |= uint16_t(1u << (act[1] - '1'));
            else return RXM_ERR_PARSE;
        }
        h.edge_kind.push_back(k);
        h.edge_sym.push_back(s);
        h.edge_to.push_back(uint16_t(to));
        h.edge_open.push_back(om);
        h.edge_close.push_back(cm);
        seen++;
    }
    if (!got_end || seen != edges) return RXM_ERR_PARSE;
    while (cur_from < states) h.edge_begin[++cur_from] = seen;

    // one block: header, then arrays (8-byte aligned pieces)
    auto pad8 = [](size_t x) { return (x + 7) & ~size_t(7); };
    const size_t o_begin = pad8(sizeof(Block));
    const size_t o_kind = o_begin + pad8(sizeof(uint32_t) * (states + 1));
    const size_t o_sym = o_kind + pad8(edges);
    const size_t o_to = o_sym + pad8(edges);
    const size_t o_open = o_to + pad8(2 * size_t(edges));
    const size_t o_close = o_open + pad8(2 * size_t(edges));
    const size_t total = o_close + pad8(2 * size_t(edges));
    char *mem = static_cast<char *>(std::calloc(1, total));
    if (!mem) return RXM_ERR_NOMEM;
    std::memcpy(mem + o_begin, h.edge_begin.data(), sizeof(uint32_t) * (states + 1));
    if (edges) {
        std::memcpy(mem + o_kind, h.edge_kind.data(), edges);
        std::memcpy(mem + o_sym, h.edge_sym.data(), edges);
        std::memcpy(mem + o_to, h.edge_to.data(), 2 * size_t(edges));
        std::memcpy(mem + o_open, h.edge_open.data(), 2 * size_t(edges));
        std::memcpy(mem + o_close, h.edge_close.data(), 2 * size_t(edges));
    }
    rxm_tables *t = reinterpret_cast<rxm_tables *>(mem);
    t->abi_version = RXM_ABI_VERSION;
    t->kind = h.kind;
    t->reversed = h.reversed ? 1 : 0;
    t->n_states = states;
    t->n_edges = edges;
    t->start = h.start;
    t->finish = h.finish;
    t->n_cells = h.n_cells;
    t->edge_begin = reinterpret_cast<const uint32_t *>(mem + o_begin);
    t->edge_kind = reinterpret_cast<const uint8_t *>(mem + o_kind);
    t->edge_sym = reinterpret_cast<const uint8_t *>(mem + o_sym);
    t->edge_to = reinterpret_cast<const uint16_t *>(mem + o_to);
    t->edge_open = reinterpret_cast<const uint16_t *>(mem + o_open);
    t->edge_close = reinterpret_cast<const uint16_t *>(mem + o_close);
    const int st = rxm_tables_validate(t);
    if (st != RXM_OK) {
        std::free(mem);
        return st == RXM_ERR_INVALID ? RXM_ERR_PARSE : st;
    }
    *out = t;
    return RXM_OK;
}

// The C entry points never throw: an allocation failure inside the C++ containers is RXM_ERR_NOMEM.
extern "C" int rxm_tables_format(const rxm_tables *t, char *buf, size_t buf_size, size_t *needed) {
    try {
        return tables_format_impl(t, buf, buf_size, needed);
    } catch (const std::bad_alloc &) {
        return RXM_ERR_NOMEM;
    } catch (...) {
        return RXM_ERR_INVALID;
    }
}
extern "C" int rxm_tables_parse(const char *text, size_t len, rxm_tables **out) {
    try {
        return tables_parse_impl(text, len, out);
    } catch (const std::bad_alloc &) {
        if (out) *out = nullptr;
        return RXM_ERR_NOMEM;
    } catch (...) {
        if (out) *out = nullptr;
        return RXM_ERR_PARSE;
    }
}

extern "C" void rxm_tables_release(rxm_tables *t) { std::free(t); }

extern "C" const char *rxm_strerror(int status) {
    switch (status) {
        case RXM_OK: return "ok";
        case RXM_ERR_INVALID: return "invalid argument or table";
        case RXM_ERR_UNSUPPORTED: return "automaton outside device limits (no CPU fallback)";
        case RXM_ERR_NO_DEVICE: return "no usable CUDA device";
        case RXM_ERR_CUDA: return "CUDA runtime error";
        case RXM_ERR_NOMEM: return "out of memory";
        case RXM_ERR_PARSE: return "malformed table text";
        case RXM_ERR_OVERFLOW: return "a string exceeded a kernel limit";
        default: return "unknown status";
    }
}
