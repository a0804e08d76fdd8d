// K4 core -- MFA, ONE THREAD per string, over the host-compiled edge programs.
//
// The same simulation as ProgSim (rxm_mfa_core.cuh) -- MFA::match / evaluateStates / evaluateState
// (mfa.cpp:215-236 / 203-213 / 136-200) with cells as spans, one slot per node and prog_stamp
// creation order -- arranged for a thread that owns its string:
//   * the thread has a POOL of a few configuration slots (words reached through base[w * STRIDE]: on the
//     device shared memory with word w of thread t at w * blockDim + t, so the 32 lanes of a warp never
//     meet in a bank whatever slot each of them touches; on the host a plain array).  The sets of the
//     simulation -- the current one, the one being built, the one the last full step started from --
//     are BIT MASKS over the pool, and a waiting configuration that a step re-inserts unchanged
//     (mfa.cpp:195-197) is not copied: its slot simply joins the new set;
//   * a configuration only walks the items that can act for it: the LEAF items of its program when
//     it is ACTIVE (first == i, mfa.cpp:161-193), the ENTER items that can insert when it WAITS or
//     sits at the end of the input (mfa.cpp:138-140, 195-197) -- lists made by the planner
//     (MfaProgram::sel);
//   * backreference blocks (mfa.cpp:176-193) are compared 8 bytes per iteration from aligned words
//     at any alignment of the two spans; a block already compared in this step is not compared again;
//   * REPEATED STEPS (ProgSim::replay, proven out on the host in round 1): when the set a step starts
//     from is the set the previous step started from, moved on by the distance delta between the two,
//     and the input it looks at equals the input the previous step looked at, its result is the
//     previous result moved on.  The thread then only checks that the input stays PERIODIC with period
//     delta, one aligned word per iteration of a flat loop (phase A below); programs and slots are not
//     touched.  On the reference's example 5 this answers more than nine steps in ten;
//   * idle steps (every configuration waiting inside a block) are skipped as in ProgSim.
// A round of a string is pre() -> phase_a() or phase_b(); the kernel (rxm_k4.cu) lets the lanes of a
// warp run the phase most of them want.  Everything is RXM_HD: tests/hostsim runs this very code on
// the CPU against the golden vectors.
#ifndef RXM_K4_CORE_CUH
#define RXM_K4_CORE_CUH

#include "rxm_mfa_core.cuh"

#ifndef K4_STAT
#define K4_STAT(x)  // host statistics (tools), compiled out
#endif

namespace rxm {

struct K4Prog {  // the edge programs and the per-key item lists (MfaProgram), the lists packed into ONE array
    const ProgItem *items;
    // [begin | count | lbeg | lcnt] (n_keys words each), [cbeg | ccnt] (n_keys * n_classes words each), then the
    // byte -> class table (256 bytes): one base pointer instead of seven -- the kernel runs at its register limit
    //   begin[key]  first item, 0xffffffff: a (node, cells) pair the host analysis did not reach
    //   count[key]  items | kProgStable
    //   lbeg[key]   first entry of the key's lists in sel;  lcnt[key]  leaves | enters << 16
    //   cbeg / ccnt[key * n_classes + class]  the leaves that can act on a letter of that class (n_classes >= 1)
    const uint32_t *lists;
    const uint16_t *sel;    // item indices relative to begin[key]
    uint32_t n_cells, n_classes, n_keys;
    RXM_HD uint32_t word(uint32_t idx) const { return lists[idx]; }
    RXM_HD uint32_t sel_at(uint32_t idx) const { return sel[idx]; }
    RXM_HD uint32_t cls(uint32_t byte) const {
        return reinterpret_cast<const uint8_t *>(lists + (4u + 2u * n_classes) * n_keys)[byte];
    }
    RXM_HD uint32_t begin(uint32_t key) const { return word(key); }
    RXM_HD uint32_t count(uint32_t key) const { return word(n_keys + key); }
    RXM_HD uint32_t lbeg(uint32_t key) const { return word(2u * n_keys + key); }
    RXM_HD uint32_t lcnt(uint32_t key) const { return word(3u * n_keys + key); }
    RXM_HD uint32_t cbeg(uint32_t key, uint32_t c) const { return word(4u * n_keys + key * n_classes + c); }
    RXM_HD uint32_t ccnt(uint32_t key, uint32_t c) const { return word((4u + n_classes) * n_keys + key * n_classes + c); }
};
// words of K4Prog::lists
RXM_HD constexpr uint32_t k4_list_words(uint32_t n_keys, uint32_t n_classes) { return (4u + 2u * n_classes) * n_keys + 64u; }

#ifndef RXM_K4_A_WORDS  // (tuning builds set it)
#define RXM_K4_A_WORDS 2  // 4: config 3 +2 %, the ten README examples +3 % (code size); 1: examples 1 and 4 +50 %
#endif
constexpr int K4_A_WORDS = RXM_K4_A_WORDS;  // aligned 8-byte words a thread checks per iteration of its own repeated-step loop
constexpr uint32_t K4_LOGN = 4;      // distinct block compares remembered per step
constexpr uint32_t K4_BURST = 64;    // iterations of the repeated-step loop per round
constexpr uint32_t K4_POOL_MAX = 32;  // slots a thread can have at most (the set masks are 32 bits)

// slots for an automaton of n_states nodes: 8 hold the current set and the one being built for every forward
// automaton of the reference's examples; the reversed ones keep more configurations waiting
RXM_HD constexpr uint32_t k4_pool_for(uint32_t n_states) { return n_states < 8u ? 8u : (n_states > 15u ? 15u : n_states); }
constexpr uint32_t K4_MAP_STATES = 16;  // up to this many nodes the new set's slot of a node is kept in a 64-bit map (4 bits each)
constexpr uint32_t K4_MAX_STATES = 40;  // larger automata go to K3 (one warp per string) as a whole
RXM_HD constexpr uint32_t k4_slot_words(uint32_t nc) { return 3u + 2u * nc; }
RXM_HD constexpr uint32_t k4_words(uint32_t nc, uint32_t pool) { return pool * k4_slot_words(nc) + 2u * K4_LOGN; }

RXM_HD int k4_ffs(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs(int(x));
#else
    return __builtin_ffs(int(x));
#endif
}
RXM_HD int k4_ffs64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x);
#else
    return __builtin_ffsll((long long)x);
#endif
}
RXM_HD int k4_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

RXM_HD uint64_t k4_ld64(const uint8_t *p) {  // p is 8-byte aligned
#if defined(__CUDA_ARCH__)
    return __ldg(reinterpret_cast<const unsigned long long *>(p));
#else
    return *reinterpret_cast<const uint64_t *>(p);
#endif
}

// L2 prefetch of the line at p (if it still holds a byte of the string): the thread's stream is read ahead of its compares
RXM_HD void k4_prefetch_l2(const uint8_t *p, const uint8_t *end) {
#if defined(__CUDA_ARCH__)
    if (p < end) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
    (void)end;
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
RXM_HD uint32_t k4_open_mask(uint32_t flags) {  // bit k <- cell k exists and is open
    uint32_t m = 0;
RXM_UNROLL
    for (int k = 0; k < NC; k++) m |= (((flags >> (3 * k)) & 3u) == 3u ? 1u : 0u) << k;
    return m;
}

template <int NC>
RXM_HD uint32_t k4_exists_mask(uint32_t flags) {  // bit k <- flags bit 3k, k < NC
    uint32_t m = 0;
RXM_UNROLL
    for (int k = 0; k < NC; k++) m |= ((flags >> (3 * k)) & 1u) << k;
    return m;
}

enum : uint32_t { K4_WANT_NONE = 0, K4_WANT_A = 1, K4_WANT_B = 2, K4_DONE = 3 };

// STRIDE: distance in words between two consecutive words of this thread (device: threads per block; host: 1)
template <int NC, int STRIDE>
struct K4Sim {
    typedef Cfg<NC> cfg_t;
    static constexpr uint32_t SW = 3u + 2u * NC;  // first | node, flags | born | start[NC] | len[NC]

    uint32_t *base;  // word 0 of this thread
    uint32_t pool;   // slots in the pool (<= K4_POOL_MAX); the compare log follows them
    bool use_map;    // the automaton has <= K4_MAP_STATES nodes and the pool <= 15 slots
    // the string in hand
    const uint8_t *s;
    uint32_t n, reversed;
    // simulation state (registers)
    uint32_t i;
    uint32_t cur, nxt, prv;  // set masks: current set / set being built / set the last full step started from
    uint32_t n_log;
    uint32_t prev_i;
    uint32_t want;        // K4_WANT_*: what the next round of this string is
    int result;
    // repeated steps
    uint32_t rp_delta;    // != 0: repeated steps of this distance are being answered
    uint32_t rp_lmax;     // the longest block the repeated step must see again (>= 1: its letter)
    uint32_t log_open;    // bit e: log entry e is a block of a cell that is still OPEN
    uint32_t unsafe;      // bit j: the new set's configuration in slot j does not move on with the step (see phase_a)
    uint64_t nmap;        // (automata of <= K4_MAP_STATES nodes) 4 bits per node: slot + 1 of the new set's configuration on it
    uint32_t rp_vp;       // verified front: s[j] == s[j - rp_delta] for every j from the first repeated step up to here
    uint32_t rp_vcap;     // the front need not pass this: the last step that may be answered looks up to here
    uint32_t rp_acc;      // distance moved on so far, not yet written to the slots
    uint64_t rp_magic;    // ceil(2^32 / rp_delta)
    bool rp_mism;         // a byte that differs stands at rp_vp
    bool have_prev, log_bad, overflow, near;
    // statistics (tests)
    uint32_t steps_run, steps_replayed;

    RXM_HD uint32_t *slotp(uint32_t j) const { return base + size_t(j) * (SW * STRIDE); }
    RXM_HD uint32_t *logp() const { return base + size_t(pool) * (SW * STRIDE); }
    RXM_HD uint8_t at(uint32_t j) const { return reversed ? s[n - 1u - j] : s[j]; }
    RXM_HD bool span_equal(uint32_t a, uint32_t b, uint32_t L) const {  // R[a, a+L) == R[b, b+L), reading direction
        if (!reversed) return k4_span_equal(s + a, s + b, L);
        return k4_span_equal(s + (n - a - L), s + (n - b - L), L);
    }

    RXM_HD void load(uint32_t j, cfg_t &c) const {
        const uint32_t *sp = slotp(j);
        c.first = sp[0];
        const uint32_t nf = sp[STRIDE];
        c.node = nf & 0xffffu;
        c.flags = nf >> 16;
        c.born = sp[2 * STRIDE];
RXM_UNROLL
        for (int k = 0; k < NC; k++) {
            c.start[k] = sp[(3 + k) * STRIDE];
            c.len[k] = sp[(3 + NC + k) * STRIDE];
        }
    }
    RXM_HD void store(uint32_t j, const cfg_t &c) const {
        uint32_t *sp = slotp(j);
        sp[0] = c.first;
        sp[STRIDE] = c.node | (c.flags << 16);
        sp[2 * STRIDE] = c.born;
RXM_UNROLL
        for (int k = 0; k < NC; k++) {
            sp[(3 + k) * STRIDE] = c.start[k];
            sp[(3 + NC + k) * STRIDE] = c.len[k];
        }
    }

    // set order of mfa.cpp:206 for two configurations on one node: (first, lowest cell name, creation)
    RXM_HD static bool key_less(uint32_t fa, uint32_t fla, uint32_t ba, uint32_t fb, uint32_t flb, uint32_t bb) {
        if (fa != fb) return fa < fb;
        // lowest cell name present: the lowest `exists` bit (flag bit 3k for cell k); no bit = empty memory, smallest
        const uint32_t ea = fla & 0x1249249u, eb = flb & 0x1249249u;
        const uint32_t la = ea & (0u - ea), lb = eb & (0u - eb);
        if (la != lb) return la < lb;
        return la != 0 && ba < bb;  // both memories empty: equal keys, the one already there stays
    }

    // the new set's slot on `node`, or K4_POOL_MAX
    RXM_HD uint32_t find_new(uint32_t node) const {
        if (use_map) {
            const uint32_t e = uint32_t(nmap >> (4u * node)) & 15u;
            return e ? e - 1u : K4_POOL_MAX;
        }
        for (uint32_t m = nxt; m; m &= m - 1u) {
            const uint32_t j = uint32_t(k4_ffs(m)) - 1u;
            K4_STAT(g_find++);
            if ((slotp(j)[STRIDE] & 0xffffu) == node) return j;
        }
        return K4_POOL_MAX;
    }
    RXM_HD uint32_t alloc() {
        const uint32_t free = ~(cur | nxt) & (pool >= 32u ? 0xffffffffu : ((1u << pool) - 1u));
        if (!free) {
            overflow = true;
            return K4_POOL_MAX;
        }
        return uint32_t(k4_ffs(free)) - 1u;
    }

    // new_states.insert, reduced on the fly to the set-minimum per node (mfa.cpp:206-211)
    // (bad: the candidate is not the previous step's candidate moved on -- see phase_a; it may lose here, not stand)
    RXM_HD void insert(const cfg_t &c, bool bad) {
        K4_STAT(g_ins++);
        if (c.first < i + 2u) near = true;  // its node will hold a configuration with first <= i + 1
        uint32_t j = find_new(c.node);
        if (j != K4_POOL_MAX) {
            const uint32_t *sp = slotp(j);
            if (!key_less(c.first, c.flags, c.born, sp[0], sp[STRIDE] >> 16, sp[2 * STRIDE])) return;
            if (!((cur >> j) & 1u)) {  // not shared with the current set: overwritten in place
                store(j, c);
                unsafe = (unsafe & ~(1u << j)) | (bad ? 1u << j : 0u);
                return;
            }
            nxt &= ~(1u << j);
        }
        j = alloc();
        if (j == K4_POOL_MAX) return;
        store(j, c);
        nxt |= 1u << j;
        unsafe = (unsafe & ~(1u << j)) | (bad ? 1u << j : 0u);
        if (use_map) nmap = (nmap & ~(15ull << (4u * c.node))) | (uint64_t(j + 1u) << (4u * c.node));
    }
    // the re-insertion of the configuration in slot js itself (mfa.cpp:195-197 / 138-140 on the root call):
    // older than anything created in this step (stamp 0); its slot joins the new set, nothing is copied
    RXM_HD void reinsert(uint32_t js, const cfg_t &root) {
        K4_STAT(g_reins++);
        if (root.first < i + 2u) near = true;
        const uint32_t j = find_new(root.node);
        if (j != K4_POOL_MAX) {
            const uint32_t *sp = slotp(j);
            if (!key_less(root.first, root.flags, 0u, sp[0], sp[STRIDE] >> 16, sp[2 * STRIDE])) return;
            nxt &= ~(1u << j);  // (a slot of the new set alone: free again)
        }
        slotp(js)[2 * STRIDE] = 0u;
        nxt |= 1u << js;
        unsafe &= ~(1u << js);
        if (use_map) nmap = (nmap & ~(15ull << (4u * root.node))) | (uint64_t(js + 1u) << (4u * root.node));
    }

    RXM_HD static uint32_t need_of(uint32_t flags, const uint32_t *len) {  // is_siffix_long_enough, mfa.cpp:116-133
        uint32_t need = 0;
RXM_UNROLL
        for (int k = 0; k < NC; k++) {
            const uint32_t fl = (flags >> (3 * k)) & 7u;
            if ((fl & 1u) && ((fl & 2u) || !(fl & 4u))) need += len[k];
        }
        return need;
    }

    // one block compare of the step at i; a span already compared in this step is answered from the log
    // (open: the cell is still open -- its text, and with it this block, grows with every step)
    RXM_HD bool compare_logged(uint32_t vs, uint32_t L, bool open) {
        uint32_t *lg = logp();
        for (uint32_t e = 0; e < n_log; e++)
            if (lg[(2 * e) * STRIDE] == vs && (lg[(2 * e + 1) * STRIDE] & 0x7fffffffu) == L) {
                if ((((log_open >> e) & 1u) != 0u) != open) log_bad = true;  // one span, two cells that part ways
                return (lg[(2 * e + 1) * STRIDE] >> 31) != 0u;
            }
        K4_STAT(g_cmp++; g_cmpbytes += L);
        const bool eq = span_equal(vs, i, L);
        if (n_log < K4_LOGN) {
            lg[(2 * n_log) * STRIDE] = vs;
            lg[(2 * n_log + 1) * STRIDE] = L | (eq ? 0x80000000u : 0u);
            if (open) log_open |= 1u << n_log;
            n_log++;
        } else {
            log_bad = true;
        }
        return eq;
    }

    // the current set == the set the previous step started from, every `first` moved on by delta and every
    // open cell grown by delta (the letters read in between).  (Creation stamps are not compared: a step
    // never reads the stamps of the set it starts from.)
    RXM_HD bool moved_on(uint32_t delta) const {
        if (k4_popc(cur) != k4_popc(prv)) return false;
        for (uint32_t mc = cur; mc; mc &= mc - 1u) {
            const uint32_t jc = uint32_t(k4_ffs(mc)) - 1u;
            if ((prv >> jc) & 1u) return false;  // the very same configuration: it has not moved
            cfg_t b;
            load(jc, b);
            bool found = false;
            for (uint32_t mp = prv; mp; mp &= mp - 1u) {
                const uint32_t jp = uint32_t(k4_ffs(mp)) - 1u;
                if ((slotp(jp)[STRIDE] & 0xffffu) != b.node) continue;
                cfg_t mv;
                load(jp, mv);
                mv.first += delta;
RXM_UNROLL
                for (int k = 0; k < NC; k++)
                    if (fl_exists(mv.flags, k) && fl_open(mv.flags, k)) {
                        if (mv.len[k] == 0) mv.start[k] = b.start[k];
                        mv.len[k] += delta;
                    }
                found = cfg_same<NC>(mv, b);
                break;
            }
            if (!found) return false;
        }
        return true;
    }

    RXM_HD void start(const uint8_t *str, uint32_t len, uint32_t rev, uint32_t start_node) {
        s = str;
        n = len;
        reversed = rev;
        i = 0;
        cur = 1u;
        nxt = prv = 0u;
        n_log = 0;
        prev_i = 0;
        want = K4_WANT_NONE;
        result = 0;
        rp_delta = rp_lmax = rp_vp = rp_vcap = rp_acc = log_open = unsafe = 0;
        nmap = 0;
        rp_magic = 0;
        rp_mism = false;
        have_prev = log_bad = overflow = near = false;
        steps_run = steps_replayed = 0;
        cfg_t c0;
        c0.first = 0;
        c0.born = 0;
        c0.flags = 0;
        c0.node = start_node;
RXM_UNROLL
        for (int k = 0; k < NC; k++) {
            c0.start[k] = 0;
            c0.len[k] = 0;
        }
        store(0, c0);
    }

    // ---- what is the next round?  (dead set: done; the set has moved on from the previous step's: A; else B) ----
    RXM_HD void pre() {
        if (want != K4_WANT_NONE) return;
        if (i < n && cur == 0u) {  // :224-225 -- and the pass at i == n runs on the empty set
            result = 0;
            want = K4_DONE;
            return;
        }
        want = K4_WANT_B;
        if (have_prev && !reversed && !log_bad && i < n && i > prev_i && moved_on(i - prev_i)) {
            const uint32_t delta = i - prev_i;
            const uint32_t *lg = logp();
            uint32_t lmax = 1;  // the letter at i
            for (uint32_t e = 0; e < n_log; e++) {
                const uint32_t lw = lg[(2 * e + 1) * STRIDE];
                // a block of an OPEN cell that matched made a candidate that lost (phase_b): whether it matches again
                // does not matter.  Every other block must be seen again as it was.
                if (((log_open >> e) & 1u) && (lw >> 31)) continue;
                const uint32_t L = lw & 0x7fffffffu;
                lmax = L > lmax ? L : lmax;
            }
            uint32_t maxf = 0;
            for (uint32_t m = cur; m; m &= m - 1u) {
                const uint32_t f = slotp(uint32_t(k4_ffs(m)) - 1u)[0];
                maxf = f > maxf ? f : maxf;
            }
            // the step at p is only answered if p + delta + 2 < n and no `first` (moved on to p) reaches n - delta:
            // everything the step and the jump after it ask about the end of the string stays as it was
            uint32_t d = maxf > i ? maxf - i : 0u;
            d = d < 2u ? 2u : d;
            if (uint64_t(i) + delta + d < n) {
                const uint32_t p_end = n - delta - d;  // steps at p < p_end may be answered
                const uint64_t vcap = uint64_t(p_end) - 1u + lmax;  // where the last of them looks up to
                rp_delta = delta;
                rp_lmax = lmax;
                rp_vp = i;
                rp_vcap = vcap < n ? uint32_t(vcap) : n;  // (a block that does not fit the rest is never answered)
                rp_acc = 0;
                rp_mism = false;
                rp_magic = 0xffffffffull / delta + 1ull;
                want = K4_WANT_A;
            }
        }
    }

    // ---- PHASE A: repeated steps ------------------------------------------------------------------
    // sigma(c) = configuration c moved on by delta: first + delta, every OPEN cell longer by delta (an open cell
    // ends where its configuration reads: start + len == first), closed cells as they are.  Let the step at
    // p = i - delta have turned the set Q into the set P in hand, and P == sigma(Q) (moved_on).  The step at i
    // then walks the same programs from sources with the same flags; if the letter is the same and every block
    // compare comes out as before, the same items fire, and a candidate is sigma of the previous step's
    // candidate -- first + 1 or + L, open cells appended to, closed cells untouched -- UNLESS its item
    //   * opens a cell (open action, or a cell created open on the way down: anchored at i, not at i - delta),
    //   * closes a cell that is open (what it freezes is delta longer than before), or
    //   * reads a cell that is still open (the block, hence the candidate's first, is delta longer).
    // Such a candidate's key never moves by less than delta, so where it LOST the per-node minimum to a
    // candidate that does move on it loses again; phase_b refuses the repetition only if one of them STANDS in
    // the result.  Then step(sigma(Q)) == sigma(step(Q)), and by induction for every following step, as long as
    //     s[j] == s[j - delta]   for j in [i, i + lmax),   lmax = the longest block to be seen again (>= 1: the letter)
    // -- equal windows give equal outcomes for blocks of closed cells, and a block of an open cell that did
    // NOT match keeps its differing letter (one that matched made a losing candidate: it does not matter).
    // (This is ProgSim::replay with its conditions made sufficient: the host study of round 1 refused every
    // step that reads an open cell -- the reference's examples 1, 3, 4 and 8 never repeated -- and did not look
    // at open / close actions.)  One iteration of the FLAT loop verifies aligned words ahead of the front rp_vp
    // and then answers every step the front has passed -- the same instructions for every lane whatever its
    // distance, block length and position.
    RXM_HD void phase_a() {
        const uint32_t delta = rp_delta;
        for (uint32_t r = 0; r < K4_BURST; r++) {
            const bool can = !rp_mism && rp_vp < rp_vcap;
            if (can) {
                // K4_A_WORDS aligned words of the input from the front on against the input delta bytes before them; their
                // loads are issued together (one memory round trip per 8 * K4_A_WORDS bytes), and only words that hold a
                // byte of the string are read
                const uint8_t *b = s + rp_vp, *end = s + n;
                const uint32_t ob = uint32_t(reinterpret_cast<uintptr_t>(b) & 7u);
                const uint8_t *b0 = b - ob;
                const uint8_t *ap = b0 - delta;
                const uint32_t oa = uint32_t(reinterpret_cast<uintptr_t>(ap) & 7u);
                const uint8_t *a0 = ap - oa;
                const uint32_t sh = oa * 8u;
                uint64_t w[K4_A_WORDS], a[K4_A_WORDS + 1];
RXM_UNROLL
                for (int u = 0; u < K4_A_WORDS; u++) w[u] = (b0 + 8 * u < end) ? k4_ld64(b0 + 8 * u) : 0ull;
RXM_UNROLL
                for (int u = 0; u < K4_A_WORDS + 1; u++)  // word u of the past: read iff it holds a byte of [b - delta, end - delta)
                    a[u] = (a0 + 8 * u + 8 > b - delta && a0 + 8 * u < end - delta) ? k4_ld64(a0 + 8 * u) : 0ull;
                k4_prefetch_l2(b0 + 512, end);
                uint32_t nvp = uint32_t((b0 + 8 * K4_A_WORDS < end ? b0 + 8 * K4_A_WORDS : end) - s);
                bool mism = false;
RXM_UNROLL
                for (int u = K4_A_WORDS - 1; u >= 0; u--) {  // the FIRST word that differs decides: walk backwards, keep the earliest
                    uint64_t x = (sh ? ((a[u] >> sh) | (a[u + 1] << (64u - sh))) : a[u]) ^ w[u];
                    if (u == 0) x &= ~0ull << (8u * ob);
                    const uint8_t *ub = b0 + 8 * u;
                    if (ub >= end) x = 0;
                    else if (ub + 8 > end) x &= (1ull << (8u * uint32_t(end - ub))) - 1ull;
                    if (x) {
                        nvp = uint32_t(ub - s) + (uint32_t(k4_ffs64(x)) - 1u) / 8u;
                        mism = true;
                    }
                }
                rp_vp = nvp;
                rp_mism = mism;
            }
            const uint32_t front = rp_vp < rp_vcap ? rp_vp : rp_vcap;
            const uint32_t target = i + rp_lmax;
            if (front >= target) {  // the steps at i, i + delta, ... whose blocks end at or before the front
                uint32_t xx = front - target;
                uint32_t k;
                if (delta >= 4096u) {
                    k = 1u + (xx >= delta ? 1u : 0u);
                } else {
                    xx = xx < (1u << 19) ? xx : (1u << 19);  // xx * delta < 2^32: the quotient below is exact
                    k = 1u + uint32_t((xx * rp_magic) >> 32);
                }
                rp_acc += k * delta;
                i += k * delta;
                steps_replayed += k;
            } else if (!can) {  // the front stands (a difference, or the end of what may be answered)
                rp_delta = 0;
                break;
            }
        }
        if (rp_delta != 0) return;  // more of it in the next round
        if (rp_acc) {               // leaving: the previous result, moved on
            const uint32_t acc = rp_acc;
            for (uint32_t m = cur; m; m &= m - 1u) {
                uint32_t *sp = slotp(uint32_t(k4_ffs(m)) - 1u);
                sp[0] += acc;
                const uint32_t fl = sp[STRIDE] >> 16;
RXM_UNROLL
                for (int k = 0; k < NC; k++)
                    if (fl_exists(fl, k) && fl_open(fl, k)) sp[(3 + NC + k) * STRIDE] += acc;
            }
            rp_acc = 0;
        }
        want = K4_WANT_B;  // the step at i is run in full
    }

    // ---- PHASE B: one step in full -- evaluateStates (mfa.cpp:203-213) over evaluateState (:136-200)
    // -- and the jump over the idle steps behind it.  The walk over (configuration, item of its list) is one
    // flat loop: the lanes of a warp meet in the same body whatever configuration each of them is on.
    // Ends with want = K4_DONE (result = 0 / 1, or 2 if a limit was met -- never a guess) or K4_WANT_NONE.
    RXM_HD void phase_b(const MfaView &t, const K4Prog &p) {
        n_log = 0;
        log_open = 0;
        unsafe = 0;
        log_bad = false;
        near = false;
        prv = 0u;  // the set before this one is not needed any more
        nxt = 0u;
        nmap = 0;
        {
            const uint32_t ch = (i < n) ? at(i) : 0u;
            const uint32_t digit_bit = (ch >= '1' && ch <= '9') ? (1u << (ch - '1')) : 0u;
            const uint32_t cc = p.cls(ch);  // an active configuration walks the leaves of this letter's class
            uint32_t todo = cur, js = 0, q = 0, nq = 0, lbase = 0, pb = 0;
            bool active = false, waiting = false, fin = false;
            cfg_t root;
            for (;;) {  // any order of the configurations: see prog_stamp
                if (q == nq) {  // next configuration of the current set
                    if (!todo) break;
                    js = uint32_t(k4_ffs(todo)) - 1u;
                    todo &= todo - 1u;
                    load(js, root);
                    q = nq = 0;
                    fin = (root.first == n);
                    active = (i != n && i == root.first);
                    waiting = (i != n && i < root.first);
                    K4_STAT(g_src++; g_active += active; g_waiting += (waiting || fin) && !active);
                    if (!(active || waiting || fin)) continue;  // behind the step: no branch of mfa.cpp:161-197 fires
                    if (!(root.node == t.finish && fin) && t.reversed && need_of(root.flags, root.len) > n - i) continue;  // :141
                    const uint32_t key = (root.node << p.n_cells) | k4_exists_mask<NC>(root.flags);
                    pb = p.begin(key);
                    if (pb == 0xffffffffu) {  // a (node, cells) pair the host analysis did not reach
                        overflow = true;
                        continue;
                    }
                    if (active) {  // LEAF items act for an active configuration: those of the letter's class
                        lbase = p.cbeg(key, cc);
                        nq = p.ccnt(key, cc);
                    } else {       // ENTER items for a waiting / final one
                        const uint32_t lc = p.lcnt(key);
                        lbase = p.lbeg(key) + (lc & 0xffffu);
                        nq = lc >> 16;
                    }
                    continue;
                }
                const uint32_t x = p.sel_at(lbase + q);
                q++;
                K4_STAT(g_items++);
                const ProgItem it = p.items[pb + x];
                if (active) {
                    const uint32_t kind = pi_kind(it), rc = pi_read_cell(it);
                    uint32_t L = 1;
                    bool fire = false, grows = false;
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
                        grows = !fresh && (fl & 2u);  // the text of an open cell grows from step to step
                        if (n - i >= L) fire = (L == 0) || compare_logged(vs, L, grows);
                    }
                    if (fire) {
                        K4_STAT(g_fire++);
                        // does the candidate move on with the step?  (see phase_a)
                        const bool bad = grows || (pi_open(it) | pi_created_open(it)) != 0u ||
                                         (pi_close(it) & k4_open_mask<NC>(root.flags)) != 0u;
                        cfg_t nx;
                        prog_working<NC>(nx, root, pi_created(it), pi_created_open(it), pi_prior_reads(it) & ~digit_bit);
                        nx.node = pi_node(it);
                        nx.born = prog_stamp(false, root.node, x);
                        nx.first += L;
                        apply_actions<NC>(nx, pi_open(it), pi_close(it), i, L);
                        insert(nx, bad);
                    }
                } else {
                    if (fin && pi_skip_final(it)) continue;  // below a call that returned at mfa.cpp:138-140
                    const uint32_t v = pi_node(it);
                    if ((v == t.finish && fin) || (waiting && pi_has_leaf(it))) {  // :138-140 / :195-197
                        if (x == 0) {  // the root call: the configuration itself
                            reinsert(js, root);
                        } else {
                            cfg_t w;
                            prog_working<NC>(w, root, pi_created(it), pi_created_open(it), 0u);
                            w.node = v;
                            w.born = prog_stamp(true, root.node, x);
                            insert(w, pi_created_open(it) != 0u);
                        }
                    }
                }
            }
        }
        steps_run++;
        have_prev = true;
        prev_i = i;
        want = K4_WANT_NONE;
        if (nxt & unsafe) log_bad = true;  // a configuration that does not move on with the step stands in the result
        if (overflow) {
            result = 2;
            want = K4_DONE;
            return;
        }
        const uint32_t before = cur;  // the set this step started from
        prv = before;
        cur = nxt;                    // states = new_states (:212)
        nxt = 0u;
        if (i == n) {  // the pass at i == n is the last (:227-228); :230-235
            result = 0;
            for (uint32_t m = cur; m; m &= m - 1u)
                if ((slotp(uint32_t(k4_ffs(m)) - 1u)[STRIDE] & 0xffffu) == t.finish) result = 1;
            want = K4_DONE;
            return;
        }
        // ---- idle steps are not run (ProgSim::run); a jump needs every `first` >= i + 2 ----
        if (!near && cur != 0u && i + 1 < n) {
            // (a) every configuration waits and is reproduced unchanged by a step (kProgStable): the steps up to
            //     the first activation / reversed-mode pruning are the identity
            bool stable = true;
            uint32_t ev = n;
            for (uint32_t m = cur; m && stable; m &= m - 1u) {
                const uint32_t j = uint32_t(k4_ffs(m)) - 1u;
                const uint32_t *sp = slotp(j);
                const uint32_t f = sp[0], nf = sp[STRIDE];
                if (f < i + 2 || f == n) stable = false;
                const uint32_t key = ((nf & 0xffffu) << p.n_cells) | k4_exists_mask<NC>(nf >> 16);
                if (p.begin(key) == 0xffffffffu || !(p.count(key) & kProgStable)) stable = false;
                if (f < ev) ev = f;
                if (t.reversed) {
                    cfg_t c;
                    load(j, c);
                    const uint32_t need = need_of(c.flags, c.len);
                    const uint32_t ps = need > n ? 0u : n - need + 1u;  // fresh: not yet tested against mfa.cpp:141
                    if (ps < ev) ev = ps;
                }
            }
            if (stable) {
                if (ev > i + 1) i = ev - 1;  // the increment below makes the next step ev
            } else if (k4_popc(cur) == k4_popc(before)) {
                // (b) no configuration was active in the step just run and it reproduced its input set:
                //     every further step does the same, bit for bit (prog_stamp does not depend on i)
                bool idle = true;
                ev = n;
                for (uint32_t m = cur; m && idle; m &= m - 1u) {
                    const uint32_t j = uint32_t(k4_ffs(m)) - 1u;
                    cfg_t c;
                    load(j, c);
                    if (c.first <= i) idle = false;
                    if (c.first < ev) ev = c.first;
                    if (t.reversed && !(c.node == t.finish && c.first == n)) {
                        const uint32_t ps = n - need_of(c.flags, c.len) + 1u;  // need <= n - i here
                        if (ps < ev) ev = ps;
                    }
                    if ((before >> j) & 1u) continue;  // re-inserted as it was: its slot is in both sets
                    bool found = false;
                    for (uint32_t mp = before; mp; mp &= mp - 1u) {
                        const uint32_t jp = uint32_t(k4_ffs(mp)) - 1u;
                        if ((slotp(jp)[STRIDE] & 0xffffu) != c.node) continue;
                        cfg_t pv;
                        load(jp, pv);
                        found = cfg_same<NC>(pv, c);
                        break;
                    }
                    if (!found) idle = false;
                }
                if (idle && ev > i + 1) i = ev - 1;
            }
        }
        i++;
    }

    // MFA::match (mfa.cpp:215-236) for one string, start to end (host tests; the kernel drives the phases
    // itself and lets the lanes of a warp run the phase most of them want)
    RXM_HD int run(const MfaView &t, const K4Prog &p, const uint8_t *str, uint32_t len) {
        start(str, len, t.reversed, t.start);
        for (;;) {
            pre();
            if (want == K4_DONE) break;
            if (want == K4_WANT_A) phase_a();
            else phase_b(t, p);
            if (want == K4_DONE) break;
        }
        return result;
    }
};

}  // namespace rxm
#endif
