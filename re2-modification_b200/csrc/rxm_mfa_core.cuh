// Core of the MFA simulation shared by the K2 (thread-per-string) kernels.
//
// Implements, per string, exactly the reference's
//   MFA::match / evaluateStates / evaluateState   mfa.cpp:215-236 / 203-213 / 136-200
//   doMemoryWriteActions / copy_memory            mfa.cpp:80-105 / 107-114
//   is_siffix_long_enough                         mfa.cpp:116-133
// on flat tables, with these representation choices (DESIGN.md "K2"):
//   * a memory cell is a span {exists, open, read, start, len} of the input in
//     reading direction -- its text is never stored (variable.h:8-41 only ever
//     appends consumed input to an open cell);
//   * the successor set is ONE SLOT PER NODE holding the minimum under the
//     reference's std::set<MemoryState> order, because evaluateStates expands
//     only the first state of each node (mfa.cpp:206-211);
//   * that order is (first, node, lowest cell name present, creation stamp):
//     all cells of a configuration are allocated while it is created, so the
//     reference's pointer comparison of map<string,Variable*> equals creation
//     order under an allocation-ordered heap (the canonical tie-break).
//
// The functions are __host__ __device__ so that tests/hostsim can run the very
// same code on the CPU against the oracle before it is spent GPU time on; the
// library itself only ever instantiates them inside kernels.
#ifndef RXM_MFA_CORE_CUH
#define RXM_MFA_CORE_CUH

#include <stdint.h>

#if defined(__CUDACC__)
#define RXM_HD __host__ __device__ __forceinline__
#else
#define RXM_HD inline
#endif

namespace rxm {

// ---- packed edge record ---------------------------------------------------------
// bits  0..1  kind (RXM_EDGE_*)      bits  2..9   literal byte
// bits 10..18 open mask (9 cells)    bits 19..27  close mask
// bits 32..47 target state
RXM_HD uint64_t pack_edge(uint32_t kind, uint32_t sym, uint32_t to, uint32_t open_mask,
                          uint32_t close_mask) {
    return uint64_t(kind & 3u) | (uint64_t(sym & 0xffu) << 2) | (uint64_t(open_mask & 0x1ffu) << 10) |
           (uint64_t(close_mask & 0x1ffu) << 19) | (uint64_t(to & 0xffffu) << 32);
}
RXM_HD uint32_t edge_kind(uint64_t r) { return uint32_t(r) & 3u; }
RXM_HD uint32_t edge_sym(uint64_t r) { return (uint32_t(r) >> 2) & 0xffu; }
RXM_HD uint32_t edge_open(uint64_t r) { return (uint32_t(r) >> 10) & 0x1ffu; }
RXM_HD uint32_t edge_close(uint64_t r) { return (uint32_t(r) >> 19) & 0x1ffu; }
RXM_HD uint32_t edge_to(uint64_t r) { return uint32_t(r >> 32) & 0xffffu; }

enum : uint32_t { kEdgeEps = 0, kEdgeLit = 1, kEdgeAny = 2, kEdgeNever = 3 };

struct MfaView {
    const uint16_t *edge_begin;  // [n_states + 1]
    const uint64_t *edges;       // [n_edges] packed records, CSR order
    uint32_t n_states;
    uint32_t start, finish;
    uint32_t reversed;
};

// ---- configuration ----------------------------------------------------------------
// flags: bit 3k = exists, 3k+1 = open, 3k+2 = read, for cell k (name k+1)
template <int NC>
struct Cfg {
    uint32_t first;
    uint32_t born;
    uint32_t flags;
    uint32_t node;
    uint32_t start[NC];
    uint32_t len[NC];
};

RXM_HD uint32_t fl_exists(uint32_t flags, int k) { return (flags >> (3 * k)) & 1u; }
RXM_HD uint32_t fl_open(uint32_t flags, int k) { return (flags >> (3 * k + 1)) & 1u; }
RXM_HD uint32_t fl_read(uint32_t flags, int k) { return (flags >> (3 * k + 2)) & 1u; }

// lowest cell name present (1..9), 0 for an empty memory
RXM_HD uint32_t lowvar(uint32_t flags) {
    const uint32_t ex = flags & 0x1249249u;  // exists bits
#if defined(__CUDA_ARCH__)
    return ex ? uint32_t(__ffs(int(ex)) - 1) / 3u + 1u : 0u;
#else
    if (!ex) return 0;
    uint32_t k = 0;
    while (!((ex >> (3 * k)) & 1u)) k++;
    return k + 1;
#endif
}

// set order of two configurations on the same node; true if a < b
template <int NC>
RXM_HD bool cfg_less(const Cfg<NC> &a, const Cfg<NC> &b) {
    if (a.first != b.first) return a.first < b.first;
    const uint32_t la = lowvar(a.flags), lb = lowvar(b.flags);
    if (la != lb) return la < lb;
    if (la == 0) return false;  // both memories empty: equal keys
    return a.born < b.born;
}

// Input in reading direction (mfa.cpp:163-166, 181-186).
struct Reader {
    const uint8_t *s;
    uint32_t n;
    uint32_t reversed;
    RXM_HD uint8_t at(uint32_t j) const { return reversed ? s[n - 1 - j] : s[j]; }
    // R[a .. a+L) == R[b .. b+L) ?
    RXM_HD bool span_equal(uint32_t a, uint32_t b, uint32_t L) const {
        if (a == b) return true;
        if (!reversed) {
            const uint8_t *p = s + a, *q = s + b;
            for (uint32_t j = 0; j < L; j++)
                if (p[j] != q[j]) return false;
        } else {
            const uint8_t *p = s + (n - a - L), *q = s + (n - b - L);
            for (uint32_t j = 0; j < L; j++)
                if (p[j] != q[j]) return false;
        }
        return true;
    }
};

// doMemoryWriteActions (mfa.cpp:80-105); consumed text = reading span [tstart, tstart+tlen)
template <int NC>
RXM_HD void apply_actions(Cfg<NC> &m, uint32_t open_mask, uint32_t close_mask, uint32_t tstart,
                          uint32_t tlen) {
#pragma unroll
    for (int k = 0; k < NC; k++) {
        const uint32_t o = (open_mask >> k) & 1u, c = (close_mask >> k) & 1u;
        uint32_t f = (m.flags >> (3 * k)) & 7u;
        if (o) {  // create if absent (:82-86), then Variable::open() + write() (:93-96)
            f = 3u;  // exists | open, read cleared
            m.start[k] = tstart;
            m.len[k] = tlen;
        } else if (f & 1u) {
            if (c) f &= ~2u;  // close (:97-99): text not appended
            else if (f & 2u) {  // open cell without an action: append (:101-103)
                if (m.len[k] == 0) m.start[k] = tstart;
                m.len[k] += tlen;
            }
        }
        m.flags = (m.flags & ~(7u << (3 * k))) | (f << (3 * k));
    }
}

// One simulation of one string.  CAP = slot capacity (>= number of states that
// can be live at once; n_states always suffices), DMAX = recursion stack depth.
// Returns 0/1, or 2 if a limit was hit (the caller reports it; never a guess).
template <int NC, int CAP, int DMAX>
struct MfaSim {
    typedef Cfg<NC> cfg_t;
    struct Frame {
        cfg_t c;
        uint32_t e, e_end;
    };

    cfg_t cur[CAP], nxt[CAP];
    Frame stack[DMAX];
    uint32_t ncur, nnxt;
    uint32_t born;
    bool overflow;

    RXM_HD void insert(const cfg_t &c) {
        for (uint32_t j = 0; j < nnxt; j++) {
            if (nxt[j].node == c.node) {
                if (cfg_less<NC>(c, nxt[j])) nxt[j] = c;
                return;
            }
        }
        if (nnxt < CAP) nxt[nnxt++] = c;
        else overflow = true;
    }

    // evaluateState (mfa.cpp:136-200) for configuration `root`, step index i
    RXM_HD void eval(const MfaView &t, const Reader &rd, const cfg_t &root, uint32_t i) {
        const uint32_t n = rd.n;
        int sp = 0;
        stack[0].c = root;
        stack[0].e = 0xffffffffu;  // "not entered yet"
        stack[0].e_end = 0;
        while (sp >= 0) {
            Frame &f = stack[sp];
            if (f.e == 0xffffffffu) {  // function entry
                if (f.c.node == t.finish && f.c.first == n) {  // :138-140
                    insert(f.c);
                    sp--;
                    continue;
                }
                if (t.reversed) {  // :141, :116-133
                    uint32_t need = 0;
#pragma unroll
                    for (int k = 0; k < NC; k++) {
                        const uint32_t fl = (f.c.flags >> (3 * k)) & 7u;
                        if ((fl & 1u) && ((fl & 2u) || !(fl & 4u))) need += f.c.len[k];
                    }
                    if (need > n - i) {
                        sp--;
                        continue;
                    }
                }
                f.e = t.edge_begin[f.c.node];
                f.e_end = t.edge_begin[f.c.node + 1];
            }
            if (f.e == f.e_end) {
                sp--;
                continue;
            }
            const uint64_t er = t.edges[f.e++];
            const uint32_t kind = edge_kind(er), sym = edge_sym(er), to = edge_to(er);
            const bool is_cell = (kind == kEdgeLit) && sym >= '1' && sym <= '9' && int(sym - '1') < NC;
            const int k = is_cell ? int(sym - '1') : 0;
            if (kind == kEdgeEps) {  // :143-147 (actions ignored)
                if (sp + 1 >= DMAX) {
                    overflow = true;
                    continue;
                }
                Frame &g = stack[sp + 1];
                g.c = f.c;
                g.c.node = to;
                g.c.born = ++born;
                g.e = 0xffffffffu;
                sp++;
            } else if (is_cell && !fl_exists(f.c.flags, k)) {  // :148-160 absent cell: epsilon-like
                if (sp + 1 >= DMAX) {
                    overflow = true;
                    continue;
                }
                Frame &g = stack[sp + 1];
                g.c = f.c;
                g.c.node = to;
                g.c.born = ++born;
                uint32_t fl = 1u;  // exists, closed, unread, empty
                if ((edge_open(er) >> k) & 1u) {
                    fl = 3u;
                    g.c.start[k] = f.c.first;
                }
                g.c.len[k] = 0;
                g.c.flags = (g.c.flags & ~(7u << (3 * k))) | (fl << (3 * k));
                g.e = 0xffffffffu;
                sp++;
            } else if (i != n && i == f.c.first) {  // :161
                const uint32_t ch = rd.at(i);
                if (kind == kEdgeAny || (kind == kEdgeLit && sym == ch)) {  // :171-175
                    cfg_t nx = f.c;  // :167 copy BEFORE any read()
                    nx.node = to;
                    nx.born = ++born;
                    apply_actions<NC>(nx, edge_open(er), edge_close(er), i, 1u);
                    nx.first += 1;
                    insert(nx);
                } else if (is_cell) {  // cell present: :176-193
                    cfg_t nx = f.c;  // copy taken before read() marks the source
                    f.c.flags |= (4u << (3 * k));  // Variable::read(): later edges inherit is_read
                    const uint32_t L = f.c.len[k], vs = f.c.start[k];
                    if (n - i >= L && rd.span_equal(vs, i, L)) {
                        nx.node = to;
                        nx.born = ++born;
                        nx.first += L;
                        apply_actions<NC>(nx, edge_open(er), edge_close(er), i, L);
                        insert(nx);
                    }
                }
            } else if (i != n && i < f.c.first) {  // :195-197 waiting inside a block
                insert(f.c);
            }
        }
    }

    // MFA::match (mfa.cpp:215-236)
    RXM_HD int run(const MfaView &t, const Reader &rd) {
        const uint32_t n = rd.n;
        overflow = false;
        born = 0;
        ncur = 1;
        cur[0].first = 0;
        cur[0].born = 0;
        cur[0].flags = 0;
        cur[0].node = t.start;
#pragma unroll
        for (int k = 0; k < NC; k++) {
            cur[0].start[k] = 0;
            cur[0].len[k] = 0;
        }
        for (uint32_t i = 0;; i++) {
            if (i < n && ncur == 0) break;  // :224-225
            nnxt = 0;
            // set order: (first, node) -- at most one configuration per node
            uint64_t done = 0;  // CAP <= 64 handled by bitmask; larger CAP by flag array below
            for (uint32_t r = 0; r < ncur; r++) {
                uint32_t best = 0xffffffffu;
                uint64_t bestkey = ~uint64_t(0);
                for (uint32_t j = 0; j < ncur; j++) {
                    if (CAP <= 64 ? ((done >> j) & 1u) : (cur[j].node & 0x80000000u)) continue;
                    const uint64_t key = (uint64_t(cur[j].first) << 32) | (cur[j].node & 0x7fffffffu);
                    if (key < bestkey) {
                        bestkey = key;
                        best = j;
                    }
                }
                if (CAP <= 64) done |= uint64_t(1) << best;
                cfg_t c = cur[best];
                if (CAP > 64) cur[best].node |= 0x80000000u;
                eval(t, rd, c, i);
            }
            // states = new_states (:212)
            ncur = nnxt;
            for (uint32_t j = 0; j < nnxt; j++) cur[j] = nxt[j];
            if (overflow) return 2;
            if (i == n) break;  // the pass at i == n is the last (:227-228)
        }
        for (uint32_t j = 0; j < ncur; j++)
            if (cur[j].node == t.finish) return 1;  // :230-235
        return 0;
    }
};

}  // namespace rxm
#endif
