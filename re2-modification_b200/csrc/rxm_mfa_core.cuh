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
#define RXM_UNROLL _Pragma("unroll")
#else
#define RXM_HD inline
#define RXM_UNROLL
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
RXM_UNROLL
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

// True if two configurations are the same state of the simulation, creation stamp aside
// (start/len of absent cells are don't-care).
template <int NC>
RXM_HD bool cfg_same(const Cfg<NC> &a, const Cfg<NC> &b) {
    if (a.first != b.first || a.node != b.node || a.flags != b.flags) return false;
RXM_UNROLL
    for (int k = 0; k < NC; k++)
        if (fl_exists(a.flags, k) && (a.len[k] != b.len[k] || (a.len[k] && a.start[k] != b.start[k]))) return false;
    return true;
}

// One simulation of one string.  CAP = slot capacity (>= number of states that
// can be live at once; n_states always suffices), DMAX = depth of the explicit
// evaluateState recursion stack.  Returns 0/1, or 2 if a limit was hit (the
// caller reports it; never a guess).
//
// evaluateState's recursion (epsilon edges, absent-cell edges) only ever changes the
// node, the creation stamp, the flag word and -- for an absent-cell edge -- the
// start/len of a cell the parent does not have.  So the depth-first walk runs on ONE
// working configuration held in registers; a level of the stack is just
// {node, stamp, flags, edge cursor}.  Restoring the parent's flags on return also undoes
// the child's read() marks, exactly like the reference's per-call copy_memory.
template <int NC, int CAP, int DMAX>
struct MfaSim {
    typedef Cfg<NC> cfg_t;

    cfg_t buf[2][CAP];
    uint32_t cnt[2];
    uint32_t st_node[DMAX], st_born[DMAX], st_flags[DMAX], st_e[DMAX];
    uint32_t born;
    uint32_t nb;  // index of the buffer being filled ("new_states")
    uint32_t steps_run, steps_skipped;  // statistics
    bool overflow;

    RXM_HD void insert(const cfg_t &c) {
        cfg_t *nxt = buf[nb];
        const uint32_t m = cnt[nb];
        for (uint32_t j = 0; j < m; j++) {
            if (nxt[j].node == c.node) {
                if (cfg_less<NC>(c, nxt[j])) nxt[j] = c;
                return;
            }
        }
        if (m < CAP) {
            nxt[m] = c;
            cnt[nb] = m + 1;
        } else {
            overflow = true;
        }
    }

    // entry of evaluateState: mfa.cpp:138-141.  false = this call returns at once
    RXM_HD bool enter(const MfaView &t, uint32_t n, const cfg_t &w, uint32_t i) {
        if (w.node == t.finish && w.first == n) {
            insert(w);
            return false;
        }
        if (t.reversed) {  // is_siffix_long_enough, mfa.cpp:116-133
            uint32_t need = 0;
RXM_UNROLL
            for (int k = 0; k < NC; k++) {
                const uint32_t fl = (w.flags >> (3 * k)) & 7u;
                if ((fl & 1u) && ((fl & 2u) || !(fl & 4u))) need += w.len[k];
            }
            if (need > n - i) return false;
        }
        return true;
    }

    // evaluateState (mfa.cpp:136-200) for configuration `root`, step index i
    RXM_HD void eval(const MfaView &t, const Reader &rd, const cfg_t &root, uint32_t i) {
        const uint32_t n = rd.n;
        cfg_t w = root;
        if (!enter(t, n, w, i)) return;
        uint32_t sp = 0;
        uint32_t e = t.edge_begin[w.node], e_end = t.edge_begin[w.node + 1];
        for (;;) {
            if (e == e_end) {  // return from this call
                if (sp == 0) return;
                sp--;
                w.node = st_node[sp];
                w.born = st_born[sp];
                w.flags = st_flags[sp];
                e = st_e[sp];
                e_end = t.edge_begin[w.node + 1];
                continue;
            }
            const uint64_t er = t.edges[e++];
            const uint32_t kind = edge_kind(er), sym = edge_sym(er), to = edge_to(er);
            const bool is_cell = (kind == kEdgeLit) && sym >= '1' && sym <= '9' && int(sym - '1') < NC;
            const int k = is_cell ? int(sym - '1') : 0;
            const bool absent = is_cell && !fl_exists(w.flags, k);
            if (kind == kEdgeEps || absent) {  // :143-147 / :148-160: recursive call on a copy
                if (sp == DMAX) {
                    overflow = true;
                    continue;
                }
                st_node[sp] = w.node;
                st_born[sp] = w.born;
                st_flags[sp] = w.flags;
                st_e[sp] = e;
                sp++;
                w.node = to;
                w.born = ++born;
                if (absent) {  // new Variable(), then open() or close() by this edge's action on it
                    uint32_t fl = 1u;
                    if ((edge_open(er) >> k) & 1u) fl = 3u;
                    w.start[k] = w.first;
                    w.len[k] = 0;
                    w.flags = (w.flags & ~(7u << (3 * k))) | (fl << (3 * k));
                }
                if (!enter(t, n, w, i)) {  // callee returned immediately
                    sp--;
                    w.node = st_node[sp];
                    w.born = st_born[sp];
                    w.flags = st_flags[sp];
                    continue;  // e, e_end still the parent's
                }
                e = t.edge_begin[w.node];
                e_end = t.edge_begin[w.node + 1];
            } else if (i != n && i == w.first) {  // :161
                const uint32_t ch = rd.at(i);
                if (kind == kEdgeAny || (kind == kEdgeLit && sym == ch)) {  // :171-175
                    cfg_t nx = w;  // :167 copy BEFORE any read()
                    nx.node = to;
                    nx.born = ++born;
                    apply_actions<NC>(nx, edge_open(er), edge_close(er), i, 1u);
                    nx.first += 1;
                    insert(nx);
                } else if (is_cell) {  // cell present: :176-193
                    const uint32_t before = w.flags;
                    w.flags |= (4u << (3 * k));  // Variable::read(): later edges inherit is_read
                    const uint32_t L = w.len[k], vs = w.start[k];
                    if (n - i >= L && rd.span_equal(vs, i, L)) {
                        cfg_t nx = w;
                        nx.flags = before;  // the copy was taken before read()
                        nx.node = to;
                        nx.born = ++born;
                        nx.first += L;
                        apply_actions<NC>(nx, edge_open(er), edge_close(er), i, L);
                        insert(nx);
                    }
                }
            } else if (i != n && i < w.first) {  // :195-197 waiting inside a block
                insert(w);
            }
        }
    }

    // one evaluateStates call (mfa.cpp:203-213): expand buf[nb^1] into buf[nb]
    RXM_HD void step(const MfaView &t, const Reader &rd, uint32_t i) {
        cfg_t *cur = buf[nb ^ 1u];
        const uint32_t ncur = cnt[nb ^ 1u];
        cnt[nb] = 0;
        // set order: (first, node) -- at most one configuration per node
        uint64_t last = 0;
        bool have_last = false;
        for (uint32_t r = 0; r < ncur; r++) {
            uint32_t best = 0;
            uint64_t bestkey = ~uint64_t(0);
            for (uint32_t j = 0; j < ncur; j++) {
                const uint64_t key = (uint64_t(cur[j].first) << 32) | cur[j].node;
                if (key < bestkey && (!have_last || key > last)) {
                    bestkey = key;
                    best = j;
                }
            }
            last = bestkey;
            have_last = true;
            eval(t, rd, cur[best], i);
        }
    }

    // MFA::match (mfa.cpp:215-236)
    RXM_HD int run(const MfaView &t, const Reader &rd) {
        const uint32_t n = rd.n;
        overflow = false;
        born = 0;
        steps_run = steps_skipped = 0;
        nb = 1;
        cnt[0] = 1;
        cnt[1] = 0;
        cfg_t &c0 = buf[0][0];
        c0.first = 0;
        c0.born = 0;
        c0.flags = 0;
        c0.node = t.start;
RXM_UNROLL
        for (int k = 0; k < NC; k++) {
            c0.start[k] = 0;
            c0.len[k] = 0;
        }
        for (uint32_t i = 0;; i++) {
            if (i < n && cnt[nb ^ 1u] == 0) break;  // :224-225
            step(t, rd, i);
            steps_run++;
            nb ^= 1u;  // states = new_states (:212): buf[nb^1] is now current, buf[nb] the previous set
            if (overflow) return 2;
            if (i == n) break;  // the pass at i == n is the last (:227-228)
            // ---- fast-forward over idle steps (exact; DESIGN.md "K2: waiting steps") ----
            // If no configuration was active in the step just run (all were waiting inside a
            // backreference block, or parked on finish with first == n) and the step reproduced
            // the set it started from, every following step does the same until the first
            // configuration becomes active or -- reversed mode -- gets pruned.  Jump to the step
            // BEFORE that event and run it normally: it hands out the creation stamps in the
            // same relative order as the step it stands for.
            const cfg_t *now = buf[nb ^ 1u];
            const cfg_t *prev = buf[nb];
            const uint32_t m = cnt[nb ^ 1u];
            if (m == 0 || m != cnt[nb] || i + 2 >= n) continue;
            uint32_t ev = n;  // next event step
            bool idle = true;
            for (uint32_t j = 0; j < m && idle; j++) {
                const cfg_t &c = now[j];
                if (c.first <= i) idle = false;  // was active in this step, or stale
                if (c.first < ev) ev = c.first;
                if (t.reversed && !(c.node == t.finish && c.first == n)) {
                    uint32_t need = 0;
RXM_UNROLL
                    for (int k = 0; k < NC; k++) {
                        const uint32_t fl = (c.flags >> (3 * k)) & 7u;
                        if ((fl & 1u) && ((fl & 2u) || !(fl & 4u))) need += c.len[k];
                    }
                    const uint32_t ps = n - need + 1;  // first step at which need > n - step (need <= n - i here)
                    if (ps < ev) ev = ps;
                }
                bool found = false;
                for (uint32_t q = 0; q < m; q++)
                    if (prev[q].node == c.node) {
                        found = cfg_same<NC>(prev[q], c);
                        break;
                    }
                if (!found) idle = false;
            }
            if (idle && ev > i + 2) {
                steps_skipped += ev - 2 - i;
                i = ev - 2;  // loop increment makes the next step ev - 1
            }
        }
        const cfg_t *fin = buf[nb ^ 1u];
        for (uint32_t j = 0; j < cnt[nb ^ 1u]; j++)
            if (fin[j].node == t.finish) return 1;  // :230-235
        return 0;
    }
};

// =====================================================================================
// Edge programs (compiled on the host by rxm_plan.cpp: compile_programs)
// =====================================================================================
struct ProgItem {
    uint32_t a;  // bit0 type (0 ENTER, 1 LEAF) | bit1 skip_if_final | bit2 has_leaf (ENTER)
                 // | bits4-5 edge kind | bits6-13 literal byte | bits16-31 node (ENTER: v, LEAF: target)
    uint32_t b;  // bits0-8 open mask | bits9-17 close mask | bits18-26 cells created on the path
    uint32_t c;  // bits0-8 created cells that were opened | bits9-17 cells marked read before this item
                 // | bits18-21 (cell index + 1) when the LEAF reads a present cell, else 0
    uint32_t d;  // reserved
};
RXM_HD bool pi_is_leaf(const ProgItem &p) { return p.a & 1u; }
RXM_HD bool pi_skip_final(const ProgItem &p) { return p.a & 2u; }
RXM_HD bool pi_has_leaf(const ProgItem &p) { return p.a & 4u; }
RXM_HD uint32_t pi_kind(const ProgItem &p) { return (p.a >> 4) & 3u; }
RXM_HD uint32_t pi_sym(const ProgItem &p) { return (p.a >> 6) & 0xffu; }
RXM_HD uint32_t pi_node(const ProgItem &p) { return p.a >> 16; }
RXM_HD uint32_t pi_open(const ProgItem &p) { return p.b & 0x1ffu; }
RXM_HD uint32_t pi_close(const ProgItem &p) { return (p.b >> 9) & 0x1ffu; }
RXM_HD uint32_t pi_created(const ProgItem &p) { return (p.b >> 18) & 0x1ffu; }
RXM_HD uint32_t pi_created_open(const ProgItem &p) { return p.c & 0x1ffu; }
RXM_HD uint32_t pi_prior_reads(const ProgItem &p) { return (p.c >> 9) & 0x1ffu; }
RXM_HD uint32_t pi_read_cell(const ProgItem &p) { return (p.c >> 18) & 0xfu; }

struct ProgView {
    const ProgItem *items;
    const uint32_t *begin;  // [node << n_cells | mask]
    const uint32_t *count;
    uint32_t n_cells;
};

RXM_HD uint32_t exists_mask(uint32_t flags) {  // bit k <- flags bit 3k
    uint32_t m = 0;
RXM_UNROLL
    for (int k = 0; k < 9; k++) m |= ((flags >> (3 * k)) & 1u) << k;
    return m;
}

// The working configuration of a program item: the root plus the cells created on the way
// down (exists, maybe open, empty, anchored at `first`) plus the is_read marks in force.
template <int NC>
RXM_HD void prog_working(Cfg<NC> &w, const Cfg<NC> &root, uint32_t created, uint32_t created_open,
                         uint32_t marks) {
    w = root;
RXM_UNROLL
    for (int k = 0; k < NC; k++) {
        if ((created >> k) & 1u) {
            const uint32_t fl = ((created_open >> k) & 1u) ? 3u : 1u;
            w.flags = (w.flags & ~(7u << (3 * k))) | (fl << (3 * k));
            w.start[k] = root.first;
            w.len[k] = 0;
        }
        if ((marks >> k) & 1u) w.flags |= 4u << (3 * k);
    }
}

// Creation order WITHOUT a counter.  Inside one step, everything that is inserted into the
// successor set is either (a) the re-insertion of the configuration that already sits on
// that node (mfa.cpp:195-197 / 138-140 on the root call) -- it was created in an earlier
// step, so it is older than anything created now -- or (b) created in this step by source
// configuration c (visited in (first, node) order) at item x of its program.  Two candidates
// only ever compete when they land on the same node with the same `first` and lowest cell:
//   * successors of LEAF items come from ACTIVE sources (first == i): all the same `first`,
//     so their visiting order is the node order;
//   * ENTER items only insert for WAITING / final sources (first > i), and their `first` is
//     the source's, so two of them tie only if the sources' `first` agree: node order again;
//   * an active source is visited before every waiting one.
// Hence (0 for (a)) or 1 + (ENTER? 1:0) << 19 | source node << 12 | item index orders exactly
// like the reference's allocation order wherever that order is consulted, and the sources
// of a step can be expanded in ANY order (K3 expands them all at once).
// ProgView::count[key] carries the item count in its low bits and this flag: a WAITING
// configuration (first > step) with this (node, cells) key is reproduced unchanged by a step --
// its program re-inserts it (the root ENTER has a leaf, mfa.cpp:195-197) and no epsilon /
// absent-cell path below it reaches another node that has a leaf.  When every configuration of a
// set is such, the steps up to the next event are the identity and need not be run at all.
constexpr uint32_t kProgStable = 1u << 30;
constexpr uint32_t kProgCountMask = kProgStable - 1u;

RXM_HD uint32_t prog_stamp(bool is_enter, uint32_t src_node, uint32_t item) {
    return 1u + ((is_enter ? 1u : 0u) << 19) + (src_node << 12) + item;
}

// Sequential interpreter of the edge programs: same results as MfaSim, no recursion.
template <int NC, int CAP>
struct ProgSim {
    typedef Cfg<NC> cfg_t;
    cfg_t buf[2][CAP];
    uint32_t cnt[2];
    uint32_t born;
    uint32_t nb;
    uint32_t steps_run, steps_skipped;
    bool overflow;
    // REPEATED STEPS (host study of round 1 that K4's phase A grew out of, DESIGN.md 6 "K4"; off unless `replay` is set): a
    // step that starts from the set the previous step started from, moved on by the distance delta
    // between the two, with the same letter and the same outcomes of the same block compares, ends
    // in the previous result moved on by delta -- so only the compares are evaluated.
    bool replay = false;
    uint32_t steps_replayed = 0;
    struct CmpLog {
        uint32_t vs, L;
        bool attempted, ok;
    };
    CmpLog cmp_log[16];
    uint32_t n_cmp = 0;
    bool log_bad = false;  // a read of an OPEN cell (its text changes) or more reads than the log holds

    RXM_HD void insert(const cfg_t &c) {
        cfg_t *nxt = buf[nb];
        const uint32_t m = cnt[nb];
        for (uint32_t j = 0; j < m; j++) {
            if (nxt[j].node == c.node) {
                if (cfg_less<NC>(c, nxt[j])) nxt[j] = c;
                return;
            }
        }
        if (m < CAP) {
            nxt[m] = c;
            cnt[nb] = m + 1;
        } else {
            overflow = true;
        }
    }

    RXM_HD void eval(const MfaView &t, const ProgView &pv, const Reader &rd, const cfg_t &root, uint32_t i) {
        const uint32_t n = rd.n;
        const bool fin = (root.first == n);
        if (!(root.node == t.finish && fin) && t.reversed) {  // mfa.cpp:141 (after the :138 test)
            uint32_t need = 0;
RXM_UNROLL
            for (int k = 0; k < NC; k++) {
                const uint32_t fl = (root.flags >> (3 * k)) & 7u;
                if ((fl & 1u) && ((fl & 2u) || !(fl & 4u))) need += root.len[k];
            }
            if (need > n - i) return;
        }
        const uint32_t key = (root.node << pv.n_cells) | (exists_mask(root.flags) & ((1u << pv.n_cells) - 1u));
        const uint32_t b = pv.begin[key];
        if (b == 0xffffffffu) {  // a (node, cells) pair the host analysis did not reach
            overflow = true;
            return;
        }
        const uint32_t cntp = pv.count[key] & kProgCountMask;
        const bool active = (i != n && i == root.first);
        const bool waiting = (i != n && i < root.first);
        const uint32_t ch = active ? rd.at(i) : 0u;
        const uint32_t digit_bit = (active && ch >= '1' && ch <= '9') ? (1u << (ch - '1')) : 0u;
        for (uint32_t x = 0; x < cntp; x++) {
            const ProgItem it = pv.items[b + x];
            if (fin && pi_skip_final(it)) continue;  // below a call that returned at mfa.cpp:138-140
            if (!pi_is_leaf(it)) {
                const uint32_t v = pi_node(it);
                const bool at_finish = (v == t.finish && fin);
                if (at_finish || (waiting && pi_has_leaf(it))) {  // :138-140 / :195-197
                    cfg_t w;
                    prog_working<NC>(w, root, pi_created(it), pi_created_open(it), 0u);
                    w.node = v;
                    w.born = (x != 0) ? prog_stamp(true, root.node, x) : 0u;
                    insert(w);
                }
                continue;
            }
            if (!active) continue;
            const uint32_t kind = pi_kind(it), sym = pi_sym(it);
            const uint32_t rc = pi_read_cell(it);
            if (kind == kEdgeAny || (kind == kEdgeLit && sym == ch)) {  // :171-175
                cfg_t nx;
                prog_working<NC>(nx, root, pi_created(it), pi_created_open(it), pi_prior_reads(it) & ~digit_bit);
                nx.node = pi_node(it);
                nx.born = prog_stamp(false, root.node, x);
                apply_actions<NC>(nx, pi_open(it), pi_close(it), i, 1u);
                nx.first += 1;
                insert(nx);
            } else if (rc) {  // :176-193, the cell is present
                const int k = int(rc) - 1;
                const bool fresh = (pi_created(it) >> k) & 1u;
                const uint32_t L = fresh ? 0u : root.len[k < NC ? k : 0];
                const uint32_t vs = root.start[k < NC ? k : 0];
                const bool attempted = (n - i >= L) && L != 0;
                const bool eq = (n - i >= L) && (L == 0 || rd.span_equal(vs, i, L));
                if (replay) {
                    if (n_cmp == 16 || (!fresh && fl_open(root.flags, k < NC ? k : 0))) log_bad = true;
                    else cmp_log[n_cmp++] = CmpLog{vs, L, attempted, eq};
                }
                if (eq) {
                    cfg_t nx;
                    prog_working<NC>(nx, root, pi_created(it), pi_created_open(it),
                                     pi_prior_reads(it) & ~digit_bit);
                    nx.node = pi_node(it);
                    nx.born = prog_stamp(false, root.node, x);
                    nx.first += L;
                    apply_actions<NC>(nx, pi_open(it), pi_close(it), i, L);
                    insert(nx);
                }
            }
        }
    }

    RXM_HD void step(const MfaView &t, const ProgView &pv, const Reader &rd, uint32_t i) {
        cfg_t *cur = buf[nb ^ 1u];
        const uint32_t ncur = cnt[nb ^ 1u];
        cnt[nb] = 0;
        for (uint32_t j = 0; j < ncur; j++) eval(t, pv, rd, cur[j], i);  // any order: see prog_stamp
    }

    // b == a with every `first` moved on by delta and every open cell grown by delta (the letters read)
    RXM_HD static bool moved_on(const cfg_t *a, uint32_t na, const cfg_t *b, uint32_t nbb, uint32_t delta) {
        if (na != nbb) return false;
        for (uint32_t j = 0; j < nbb; j++) {
            bool found = false;
            for (uint32_t q = 0; q < na; q++) {
                if (a[q].node != b[j].node) continue;
                cfg_t mv = a[q];
                mv.first += delta;
RXM_UNROLL
                for (int k = 0; k < NC; k++)
                    if (fl_exists(mv.flags, k) && fl_open(mv.flags, k)) {
                        if (mv.len[k] == 0) mv.start[k] = b[j].start[k];
                        mv.len[k] += delta;
                    }
                found = cfg_same<NC>(mv, b[j]) && mv.born == b[j].born;
                break;
            }
            if (!found) return false;
        }
        return true;
    }

    RXM_HD int run(const MfaView &t, const ProgView &pv, const Reader &rd) {
        const uint32_t n = rd.n;
        bool have_prev = false;
        uint32_t prev_i = 0;
        n_cmp = 0;
        log_bad = false;
        steps_replayed = 0;
        overflow = false;
        born = 0;
        steps_run = steps_skipped = 0;
        nb = 1;
        cnt[0] = 1;
        cnt[1] = 0;
        cfg_t &c0 = buf[0][0];
        c0.first = 0;
        c0.born = 0;
        c0.flags = 0;
        c0.node = t.start;
RXM_UNROLL
        for (int k = 0; k < NC; k++) {
            c0.start[k] = 0;
            c0.len[k] = 0;
        }
        for (uint32_t i = 0;; i++) {
            if (i < n && cnt[nb ^ 1u] == 0) break;
#ifdef RXM_PROGSIM_OBSERVE  // tests/hostsim: statistics on the sets the steps start from
            RXM_PROGSIM_OBSERVE(buf[nb ^ 1u], cnt[nb ^ 1u], i, rd);
#endif
            if (replay && have_prev && !t.reversed && !log_bad && i < n && i > prev_i &&
                moved_on(buf[nb], cnt[nb], buf[nb ^ 1u], cnt[nb ^ 1u], i - prev_i)) {
                const uint32_t delta = i - prev_i;
                cfg_t *P = buf[nb ^ 1u];
                const uint32_t m = cnt[nb ^ 1u];
                for (;;) {
                    // everything the step and the jump after it ask about the end of the string stays as it was
                    bool same = uint64_t(i) + delta + 2 < n && rd.at(i) == rd.at(prev_i);
                    for (uint32_t j = 0; j < m && same; j++)
                        if (uint64_t(P[j].first) + delta >= n) same = false;
                    for (uint32_t c = 0; c < n_cmp && same; c++) {
                        const CmpLog &cl = cmp_log[c];
                        if (!cl.attempted) continue;  // the cell's text was too long for the rest: it still is
                        if (n - i < cl.L || rd.span_equal(cl.vs, i, cl.L) != cl.ok) same = false;
                    }
                    if (!same) break;
                    for (uint32_t j = 0; j < m; j++) {  // the previous result, moved on by delta
                        P[j].first += delta;
RXM_UNROLL
                        for (int k = 0; k < NC; k++)
                            if (fl_exists(P[j].flags, k) && fl_open(P[j].flags, k)) P[j].len[k] += delta;
                    }
                    i += delta;
                    steps_replayed++;
                }
            }
            n_cmp = 0;
            log_bad = false;
            have_prev = true;
            prev_i = i;
            step(t, pv, rd, i);
            steps_run++;
            nb ^= 1u;
            if (overflow) return 2;
            if (i == n) break;
            // Fast-forward over idle steps as in MfaSim::run, but with prog_stamp the stamps do not
            // depend on the step index: once an idle step reproduces its input set, every further
            // idle step reproduces it BIT FOR BIT, so the jump goes straight to the event step.
            const cfg_t *now = buf[nb ^ 1u];
            const cfg_t *prev = buf[nb];
            const uint32_t m = cnt[nb ^ 1u];
            if (m == 0 || i + 1 >= n) continue;
            {   // every configuration waiting and stable (kProgStable): the coming steps are the
                // identity until the first activation / reversed-mode pruning -- skip them unrun
                bool stable = true;
                uint32_t evs = n;
                for (uint32_t j = 0; j < m && stable; j++) {
                    const cfg_t &c = now[j];
                    if (c.first < i + 2 || c.first == n) stable = false;
                    const uint32_t key = (c.node << pv.n_cells) | (exists_mask(c.flags) & ((1u << pv.n_cells) - 1u));
                    if (pv.begin[key] == 0xffffffffu || !(pv.count[key] & kProgStable)) stable = false;
                    if (c.first < evs) evs = c.first;
                    if (t.reversed) {
                        uint32_t need = 0;
RXM_UNROLL
                        for (int k = 0; k < NC; k++) {
                            const uint32_t fl = (c.flags >> (3 * k)) & 7u;
                            if ((fl & 1u) && ((fl & 2u) || !(fl & 4u))) need += c.len[k];
                        }
                        const uint32_t ps = need > n ? 0u : n - need + 1;  // fresh configurations: not yet tested against mfa.cpp:141
                        if (ps < evs) evs = ps;
                    }
                }
                if (stable && evs > i + 1) {
                    steps_skipped += evs - 1 - i;
                    i = evs - 1;
                    continue;
                }
            }
            if (m != cnt[nb]) continue;
            uint32_t ev = n;
            bool idle = true;
            for (uint32_t j = 0; j < m && idle; j++) {
                const cfg_t &c = now[j];
                if (c.first <= i) idle = false;
                if (c.first < ev) ev = c.first;
                if (t.reversed && !(c.node == t.finish && c.first == n)) {
                    uint32_t need = 0;
RXM_UNROLL
                    for (int k = 0; k < NC; k++) {
                        const uint32_t fl = (c.flags >> (3 * k)) & 7u;
                        if ((fl & 1u) && ((fl & 2u) || !(fl & 4u))) need += c.len[k];
                    }
                    const uint32_t ps = n - need + 1;
                    if (ps < ev) ev = ps;
                }
                bool found = false;
                for (uint32_t q = 0; q < m; q++)
                    if (prev[q].node == c.node) {
                        found = cfg_same<NC>(prev[q], c);
                        break;
                    }
                if (!found) idle = false;
            }
            if (idle && ev > i + 1) {
                steps_skipped += ev - 1 - i;
                i = ev - 1;  // the loop increment makes the next step ev
            }
        }
        const cfg_t *f = buf[nb ^ 1u];
        for (uint32_t j = 0; j < cnt[nb ^ 1u]; j++)
            if (f[j].node == t.finish) return 1;
        return 0;
    }
};

}  // namespace rxm
#endif
