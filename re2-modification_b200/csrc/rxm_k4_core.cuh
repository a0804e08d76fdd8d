// K4 core -- MFA, ONE THREAD per string, over the host-compiled edge programs.
//
// The same simulation as ProgSim (rxm_mfa_core.cuh) -- MFA::match / evaluateStates / evaluateState
// (mfa.cpp:215-236 / 203-213 / 136-200) with cells as spans, one slot per node and prog_stamp
// creation order -- arranged for a thread that owns its string:
//   * the two sets live in a few dozen words of per-thread storage reached through K4Mem (on the
//     device: shared memory, word w of thread t at w * blockDim + t, so the 32 lanes of a warp never
//     meet in a bank whatever slot each of them touches; on the host: a plain array);
//   * a configuration only walks the items that can act for it: the LEAF items of its program when
//     it is ACTIVE (first == i, mfa.cpp:161-193), the ENTER items that can insert when it WAITS or
//     sits at the end of the input (mfa.cpp:138-140, 195-197) -- lists made by the planner
//     (MfaProgram::sel);
//   * backreference blocks (mfa.cpp:176-193) are compared 8 bytes per iteration from aligned words
//     at any alignment of the two spans; a block already compared in this step is not compared again;
//   * REPEATED STEPS are answered by their block compares alone (ProgSim::replay, proven out on the
//     host in round 1): when the set a step starts from is the set the previous step started from,
//     moved on by the distance between the two, the letter is the same and every block compare of the
//     previous step has the same outcome at the new position, the result is the previous result moved
//     on -- neither programs nor slots are touched.  On the reference's example 5 this answers more
//     than nine steps in ten;
//   * idle steps (every configuration waiting inside a block) are skipped as in ProgSim.
// Everything is RXM_HD: tests/hostsim runs this very code on the CPU against the golden vectors.
#ifndef RXM_K4_CORE_CUH
#define RXM_K4_CORE_CUH

#include "rxm_mfa_core.cuh"

namespace rxm {

struct K4Prog {  // the edge programs and the per-key item lists (MfaProgram)
    const ProgItem *items;
    const uint32_t *begin;  // [key] first item, 0xffffffff: a (node, cells) pair the host analysis did not reach
    const uint32_t *count;  // [key] items | kProgStable
    const uint32_t *lbeg;   // [key] first entry of the key's lists in sel
    const uint32_t *lcnt;   // [key] leaves | enters << 16
    const uint16_t *sel;    // item indices relative to begin[key]
    uint32_t n_cells;
};

// per-thread words: word w of this thread is base[w * stride]
struct K4Mem {
    uint32_t *base;
    uint32_t stride;
    RXM_HD uint32_t &at(uint32_t w) const { return base[size_t(w) * stride]; }
};

constexpr uint32_t K4_LOGN = 4;     // distinct block compares remembered per step
constexpr uint32_t K4_BURST = 16;   // repeated steps answered before the thread looks up again

RXM_HD constexpr uint32_t k4_slot_words(uint32_t nc) { return 3u + 2u * nc; }
RXM_HD constexpr uint32_t k4_words(uint32_t nc, uint32_t maxl) { return 2u * maxl * k4_slot_words(nc) + 2u * K4_LOGN; }

RXM_HD uint64_t k4_ld64(const uint8_t *p) {  // p is 8-byte aligned
#if defined(__CUDA_ARCH__)
    return __ldg(reinterpret_cast<const unsigned long long *>(p));
#else
    return *reinterpret_cast<const uint64_t *>(p);
#endif
}

// a[0, L) == b[0, L) ?  Only aligned 8-byte words that hold at least one byte of a span are read.
RXM_HD bool k4_span_equal(const uint8_t *a, const uint8_t *b, uint32_t L) {
    if (a == b || L == 0) return true;
    const uint32_t ob = uint32_t(reinterpret_cast<uintptr_t>(b) & 7u);
    const uint8_t *b0 = b - ob;  // aligned; the words of b are the reference grid
    const uint8_t *ap = a - ob;  // the byte of a's stream that stands against b0 (up to 7 bytes before a: masked)
    const uint32_t oa = uint32_t(reinterpret_cast<uintptr_t>(ap) & 7u);
    const uint8_t *a0 = ap - oa;  // aligned
    const uint32_t sh = oa * 8u;
    const uint8_t *a_end = a + L;
    const uint32_t total = ob + L;  // bytes from b0 to the end of the span
    uint64_t lo = (a0 + 8 > a) ? k4_ld64(a0) : 0ull;  // skipped when the whole word lies before a
    uint64_t diff = 0;
    for (uint32_t k = 0; k < total; k += 8u) {
        const uint8_t *an = a0 + k + 8u;
        const uint64_t next = (an < a_end) ? k4_ld64(an) : 0ull;
        const uint64_t wa = sh ? ((lo >> sh) | (next << (64u - sh))) : lo;
        uint64_t x = wa ^ k4_ld64(b0 + k);
        if (k == 0) x &= ~0ull << (8u * ob);
        const uint32_t rem = total - k;
        if (rem < 8u) x &= (1ull << (8u * rem)) - 1ull;
        diff |= x;
        if (diff) return false;
        lo = next;
    }
    return true;
}

template <int NC>
struct K4Sim {
    typedef Cfg<NC> cfg_t;
    static constexpr uint32_t SW = 3u + 2u * NC;

    K4Mem mem;
    uint32_t maxl;
    // the string in hand
    const uint8_t *s;
    uint32_t n, reversed;
    // simulation state (registers)
    uint32_t i;
    uint32_t nb;          // buffer being filled; nb ^ 1 holds the current set
    uint32_t cnt0, cnt1;  // configurations in buffer 0 / 1
    uint32_t n_log;
    uint32_t prev_i;
    uint32_t rp_delta;    // != 0: repeated steps of this distance are being answered
    uint32_t rp_ch;
    bool have_prev, log_bad, overflow;
    // statistics (tests)
    uint32_t steps_run, steps_replayed;

    RXM_HD uint32_t cnt(uint32_t b) const { return b ? cnt1 : cnt0; }
    RXM_HD void set_cnt(uint32_t b, uint32_t v) {
        if (b) cnt1 = v;
        else cnt0 = v;
    }
    RXM_HD uint32_t slot(uint32_t b, uint32_t j) const { return (b * maxl + j) * SW; }
    RXM_HD uint32_t log0() const { return 2u * maxl * SW; }

    RXM_HD uint8_t at(uint32_t j) const { return reversed ? s[n - 1u - j] : s[j]; }
    RXM_HD bool span_equal(uint32_t a, uint32_t b, uint32_t L) const {  // R[a, a+L) == R[b, b+L), reading direction
        if (!reversed) return k4_span_equal(s + a, s + b, L);
        return k4_span_equal(s + (n - a - L), s + (n - b - L), L);
    }

    RXM_HD void load(uint32_t b, uint32_t j, cfg_t &c) const {
        const uint32_t o = slot(b, j);
        c.first = mem.at(o);
        const uint32_t nf = mem.at(o + 1);
        c.node = nf & 0xffffu;
        c.flags = nf >> 16;
        c.born = mem.at(o + 2);
RXM_UNROLL
        for (int k = 0; k < NC; k++) {
            c.start[k] = mem.at(o + 3 + k);
            c.len[k] = mem.at(o + 3 + NC + k);
        }
    }
    RXM_HD void store(uint32_t o, const cfg_t &c) const {
        mem.at(o) = c.first;
        mem.at(o + 1) = c.node | (c.flags << 16);
        mem.at(o + 2) = c.born;
RXM_UNROLL
        for (int k = 0; k < NC; k++) {
            mem.at(o + 3 + k) = c.start[k];
            mem.at(o + 3 + NC + k) = c.len[k];
        }
    }

    // new_states.insert, reduced on the fly to the set-minimum per node (mfa.cpp:206-211)
    RXM_HD void insert(const cfg_t &c) {
        const uint32_t m = cnt(nb);
        for (uint32_t j = 0; j < m; j++) {
            const uint32_t o = slot(nb, j);
            const uint32_t nf = mem.at(o + 1);
            if ((nf & 0xffffu) != c.node) continue;
            const uint32_t ef = mem.at(o);
            bool less;
            if (c.first != ef) less = c.first < ef;
            else {
                const uint32_t la = lowvar(c.flags), lb = lowvar(nf >> 16);
                if (la != lb) less = la < lb;
                else less = la != 0 && c.born < mem.at(o + 2);
            }
            if (less) store(o, c);
            return;
        }
        if (m < maxl) {
            store(slot(nb, m), c);
            set_cnt(nb, m + 1);
        } else {
            overflow = true;
        }
    }

    RXM_HD static uint32_t need_of(const cfg_t &c) {  // is_siffix_long_enough, mfa.cpp:116-133
        uint32_t need = 0;
RXM_UNROLL
        for (int k = 0; k < NC; k++) {
            const uint32_t fl = (c.flags >> (3 * k)) & 7u;
            if ((fl & 1u) && ((fl & 2u) || !(fl & 4u))) need += c.len[k];
        }
        return need;
    }

    // one block compare of the step at i; a span already compared in this step is answered from the log
    RXM_HD bool compare_logged(uint32_t vs, uint32_t L) {
        const uint32_t l0 = log0();
        for (uint32_t e = 0; e < n_log; e++)
            if (mem.at(l0 + 2 * e) == vs && (mem.at(l0 + 2 * e + 1) & 0x7fffffffu) == L)
                return (mem.at(l0 + 2 * e + 1) >> 31) != 0u;
        const bool eq = span_equal(vs, i, L);
        if (n_log < K4_LOGN) {
            mem.at(l0 + 2 * n_log) = vs;
            mem.at(l0 + 2 * n_log + 1) = L | (eq ? 0x80000000u : 0u);
            n_log++;
        } else {
            log_bad = true;
        }
        return eq;
    }

    // evaluateState (mfa.cpp:136-200) for one configuration of the current set
    RXM_HD void eval(const MfaView &t, const K4Prog &p, const cfg_t &root) {
        const bool fin = (root.first == n);
        const bool active = (i != n && i == root.first);
        const bool waiting = (i != n && i < root.first);
        if (!(active || waiting || fin)) return;  // behind the step: no branch of mfa.cpp:161-197 fires
        if (!(root.node == t.finish && fin) && t.reversed) {  // mfa.cpp:141 (after the :138 test)
            if (need_of(root) > n - i) return;
        }
        const uint32_t key = (root.node << p.n_cells) | (exists_mask(root.flags) & ((1u << p.n_cells) - 1u));
        const uint32_t pb = p.begin[key];
        if (pb == 0xffffffffu) {
            overflow = true;
            return;
        }
        const uint32_t lb = p.lbeg[key], lc = p.lcnt[key];
        if (active) {
            const uint32_t ch = at(i);
            const uint32_t digit_bit = (ch >= '1' && ch <= '9') ? (1u << (ch - '1')) : 0u;
            const uint32_t nl = lc & 0xffffu;
            for (uint32_t q = 0; q < nl; q++) {
                const uint32_t x = p.sel[lb + q];
                const ProgItem it = p.items[pb + x];
                const uint32_t kind = pi_kind(it), rc = pi_read_cell(it);
                uint32_t L = 1;
                bool fire = false;
                if (kind == kEdgeAny || (kind == kEdgeLit && pi_sym(it) == ch)) {  // :171-175
                    fire = true;
                } else if (rc) {  // :176-193, the cell is present
                    const int k = int(rc) - 1;
                    const bool fresh = (pi_created(it) >> k) & 1u;
                    uint32_t vs = 0, fl = 0;
                    L = 0;
RXM_UNROLL
                    for (int kk = 0; kk < NC; kk++)
                        if (kk == k) {
                            L = fresh ? 0u : root.len[kk];
                            vs = root.start[kk];
                            fl = (root.flags >> (3 * kk)) & 7u;
                        }
                    if (!fresh && (fl & 2u)) log_bad = true;  // the text of an open cell changes from step to step
                    if (n - i >= L) fire = (L == 0) || compare_logged(vs, L);
                }
                if (fire) {
                    cfg_t nx;
                    prog_working<NC>(nx, root, pi_created(it), pi_created_open(it), pi_prior_reads(it) & ~digit_bit);
                    nx.node = pi_node(it);
                    nx.born = prog_stamp(false, root.node, x);
                    nx.first += L;
                    apply_actions<NC>(nx, pi_open(it), pi_close(it), i, L);
                    insert(nx);
                }
            }
        } else {
            const uint32_t ne = lc >> 16;
            for (uint32_t q = 0; q < ne; q++) {
                const uint32_t x = p.sel[lb + (lc & 0xffffu) + q];
                const ProgItem it = p.items[pb + x];
                if (fin && pi_skip_final(it)) continue;  // below a call that returned at mfa.cpp:138-140
                const uint32_t v = pi_node(it);
                if ((v == t.finish && fin) || (waiting && pi_has_leaf(it))) {  // :138-140 / :195-197
                    cfg_t w;
                    prog_working<NC>(w, root, pi_created(it), pi_created_open(it), 0u);
                    w.node = v;
                    w.born = (x != 0) ? prog_stamp(true, root.node, x) : 0u;
                    insert(w);
                }
            }
        }
    }

    // the current set == the set the previous step started from, every `first` moved on by delta and every
    // open cell grown by delta (the letters read in between)
    RXM_HD bool moved_on(uint32_t delta) const {
        const uint32_t pbuf = nb, cbuf = nb ^ 1u;
        const uint32_t m = cnt(cbuf);
        if (cnt(pbuf) != m) return false;
        for (uint32_t j = 0; j < m; j++) {
            cfg_t b;
            load(cbuf, j, b);
            bool found = false;
            for (uint32_t q = 0; q < m; q++) {
                const uint32_t o = slot(pbuf, q);
                if ((mem.at(o + 1) & 0xffffu) != b.node) continue;
                cfg_t mv;
                load(pbuf, q, mv);
                mv.first += delta;
RXM_UNROLL
                for (int k = 0; k < NC; k++)
                    if (fl_exists(mv.flags, k) && fl_open(mv.flags, k)) {
                        if (mv.len[k] == 0) mv.start[k] = b.start[k];
                        mv.len[k] += delta;
                    }
                found = cfg_same<NC>(mv, b) && mv.born == b.born;
                break;
            }
            if (!found) return false;
        }
        return true;
    }

    RXM_HD void start(const uint8_t *str, uint32_t len, uint32_t rev, const MfaView &t) {
        s = str;
        n = len;
        reversed = rev;
        i = 0;
        nb = 1;
        cnt0 = 1;
        cnt1 = 0;
        n_log = 0;
        prev_i = 0;
        rp_delta = 0;
        rp_ch = 0;
        have_prev = false;
        log_bad = false;
        overflow = false;
        steps_run = steps_replayed = 0;
        cfg_t c0;
        c0.first = 0;
        c0.born = 0;
        c0.flags = 0;
        c0.node = t.start;
RXM_UNROLL
        for (int k = 0; k < NC; k++) {
            c0.start[k] = 0;
            c0.len[k] = 0;
        }
        store(slot(0, 0), c0);
    }

    // One round of MFA::match's loop (mfa.cpp:221-228): a burst of repeated steps answered by their block
    // compares, or one step run in full plus the jump over the idle steps behind it.  Returns true when
    // the string is done: result = 0 / 1, or 2 if a limit was met (never a guess).
    RXM_HD bool advance(const MfaView &t, const K4Prog &p, int &result) {
        bool generic = true;
        if (rp_delta == 0) {
            if (i < n && cnt(nb ^ 1u) == 0) {  // :224-225 -- and the pass at i == n runs on the empty set
                result = 0;
                return true;
            }
            if (have_prev && !reversed && !log_bad && i < n && i > prev_i && moved_on(i - prev_i)) {
                rp_delta = i - prev_i;
                rp_ch = at(prev_i);
            }
        }
        if (rp_delta) {
            const uint32_t delta = rp_delta, cbuf = nb ^ 1u, m = cnt(cbuf), l0 = log0();
            uint32_t maxf = 0;
            for (uint32_t j = 0; j < m; j++) {
                const uint32_t f = mem.at(slot(cbuf, j));
                maxf = f > maxf ? f : maxf;
            }
            uint32_t acc = 0;
            for (uint32_t r = 0; r < K4_BURST; r++) {
                // everything the step and the jump after it ask about the end of the string stays as it was
                bool same = uint64_t(i) + delta + 2 < n && uint64_t(maxf) + acc + delta < n && at(i) == rp_ch;
                for (uint32_t e = 0; e < n_log && same; e++) {
                    const uint32_t vs = mem.at(l0 + 2 * e), lw = mem.at(l0 + 2 * e + 1);
                    const uint32_t L = lw & 0x7fffffffu;
                    if (n - i < L || span_equal(vs, i, L) != ((lw >> 31) != 0u)) same = false;
                }
                if (!same) {
                    rp_delta = 0;
                    break;
                }
                acc += delta;
                i += delta;
                steps_replayed++;
            }
            if (acc) {  // the previous result, moved on
                for (uint32_t j = 0; j < m; j++) {
                    const uint32_t o = slot(cbuf, j);
                    mem.at(o) += acc;
                    const uint32_t fl = mem.at(o + 1) >> 16;
RXM_UNROLL
                    for (int k = 0; k < NC; k++)
                        if (fl_exists(fl, k) && fl_open(fl, k)) mem.at(o + 3 + NC + k) += acc;
                }
            }
            generic = (rp_delta == 0);
        }
        if (!generic) return false;

        // ---- one step in full: evaluateStates (mfa.cpp:203-213) ----
        n_log = 0;
        log_bad = false;
        have_prev = true;
        prev_i = i;
        {
            const uint32_t cbuf = nb ^ 1u, m = cnt(cbuf);
            set_cnt(nb, 0);
            for (uint32_t j = 0; j < m; j++) {  // any order: see prog_stamp
                cfg_t c;
                load(cbuf, j, c);
                eval(t, p, c);
            }
        }
        steps_run++;
        nb ^= 1u;  // states = new_states (:212)
        if (overflow) {
            result = 2;
            return true;
        }
        const uint32_t cbuf = nb ^ 1u, m = cnt(cbuf);
        if (i == n) {  // the pass at i == n is the last (:227-228); :230-235
            result = 0;
            for (uint32_t j = 0; j < m; j++)
                if ((mem.at(slot(cbuf, j) + 1) & 0xffffu) == t.finish) result = 1;
            return true;
        }
        // ---- idle steps are not run (ProgSim::run) ----
        if (m != 0 && i + 1 < n) {
            // (a) every configuration waits (first >= i + 2) and is reproduced unchanged by a step
            //     (kProgStable): the steps up to the first activation / reversed-mode pruning are the identity
            bool stable = true;
            uint32_t ev = n;
            for (uint32_t j = 0; j < m && stable; j++) {
                const uint32_t o = slot(cbuf, j);
                const uint32_t f = mem.at(o), nf = mem.at(o + 1);
                if (f < i + 2 || f == n) stable = false;
                const uint32_t key = ((nf & 0xffffu) << p.n_cells) | (exists_mask(nf >> 16) & ((1u << p.n_cells) - 1u));
                if (p.begin[key] == 0xffffffffu || !(p.count[key] & kProgStable)) stable = false;
                if (f < ev) ev = f;
                if (t.reversed) {
                    cfg_t c;
                    load(cbuf, j, c);
                    const uint32_t need = need_of(c);
                    const uint32_t ps = need > n ? 0u : n - need + 1u;  // fresh: not yet tested against mfa.cpp:141
                    if (ps < ev) ev = ps;
                }
            }
            if (stable) {
                if (ev > i + 1) i = ev - 1;  // the increment below makes the next step ev
            } else if (m == cnt(nb)) {
                // (b) no configuration was active in the step just run and it reproduced its input set:
                //     every further step does the same, bit for bit (prog_stamp does not depend on i)
                bool idle = true;
                ev = n;
                for (uint32_t j = 0; j < m && idle; j++) {
                    cfg_t c;
                    load(cbuf, j, c);
                    if (c.first <= i) idle = false;
                    if (c.first < ev) ev = c.first;
                    if (t.reversed && !(c.node == t.finish && c.first == n)) {
                        const uint32_t ps = n - need_of(c) + 1u;  // need <= n - i here
                        if (ps < ev) ev = ps;
                    }
                    bool found = false;
                    for (uint32_t q = 0; q < m; q++) {
                        if ((mem.at(slot(nb, q) + 1) & 0xffffu) != c.node) continue;
                        cfg_t pv;
                        load(nb, q, pv);
                        found = cfg_same<NC>(pv, c);
                        break;
                    }
                    if (!found) idle = false;
                }
                if (idle && ev > i + 1) i = ev - 1;
            }
        }
        i++;
        return false;
    }

    // MFA::match (mfa.cpp:215-236) for one string, start to end (host tests; the kernel drives advance itself)
    RXM_HD int run(const MfaView &t, const K4Prog &p, const uint8_t *str, uint32_t len) {
        start(str, len, t.reversed, t);
        int r = 0;
        while (!advance(t, p, r)) {
        }
        return r;
    }
};

}  // namespace rxm
#endif
