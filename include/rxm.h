/*
 * rxm.h -- C ABI of the B200 batch matcher ("rxm" = regex/MFA matcher).
 *
 * This is the drop-in boundary for ONE path of Danya-Is/re2-modification: the
 * per-string automaton simulation that `./diploma -match` runs once per input
 * token.  Everything before it (parse, bnf, reverse, automaton construction)
 * stays the reference's own C++ on the CPU; a flattening stage
 * (re2-modification_b200/csrc/rxm_flatten.hpp) turns the object graph that
 * Regexp::compile returns into the plain-old-data `rxm_tables` below, and the
 * entry points here replace, for whole batches,
 *
 *     bool MFA::match(string str)               automata.h:69,  mfa.cpp:215-236
 *     bool Automata::match(const string& str)   automata.h:42,  automata.cpp:177-210
 *
 * as called from the `-match` loop, matchers/match.cpp:22-31.  The reference has
 * no FFI of its own; a maintainer binds these functions
 * from match.cpp directly -- see INTEGRATION.md.
 *
 * Rules of the boundary
 *   - plain C: pointers and sizes only, no C++/torch types; never throws;
 *   - every function returns an `int` status (RXM_OK == 0) except rxm_strerror;
 *   - there is NO CPU fallback: without a usable CUDA device every compute
 *     entry point returns RXM_ERR_NO_DEVICE / RXM_ERR_CUDA;
 *   - result bits are bit-exact with the reference for the same
 *     (expression, flags, string); see DESIGN.md for the canonical tie-break.
 */
#ifndef RXM_H
#define RXM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RXM_ABI_VERSION 1u

/* ---- status codes ------------------------------------------------------- */
enum {
    RXM_OK = 0,
    RXM_ERR_INVALID = 1,      /* NULL / malformed argument or table            */
    RXM_ERR_UNSUPPORTED = 2,  /* automaton outside device limits (NOT a fallback) */
    RXM_ERR_NO_DEVICE = 3,    /* no CUDA device / bad ordinal                  */
    RXM_ERR_CUDA = 4,         /* a CUDA runtime call failed; see rxm_last_cuda_error */
    RXM_ERR_NOMEM = 5,        /* host or device allocation failed              */
    RXM_ERR_PARSE = 6,        /* rxm_tables_parse: bad text                    */
    RXM_ERR_OVERFLOW = 7      /* a string exceeded a kernel limit (reported, never guessed) */
};

/* ---- automaton tables (host, plain old data) ----------------------------- */
enum { RXM_KIND_NFA = 0, RXM_KIND_MFA = 1 };

/* Edge kinds.  `by` is the label string of Edge / MemoryEdge (edge.h:15,34-49). */
enum {
    RXM_EDGE_EPS = 0,   /* by == "" (or "ε" after MFA::makeDOTFile, mfa.cpp:40-42) */
    RXM_EDGE_LIT = 1,   /* by is one byte, compared as text with the input byte
                           (automata.cpp:111, mfa.cpp:171); edge_sym = that byte.
                           In an MFA the bytes '1'..'9' ALSO name a memory cell
                           (mfa.cpp:148,176) -- the kernels apply both meanings
                           in the reference's order.                              */
    RXM_EDGE_ANY = 2,   /* by == "."  : matches every input byte                  */
    RXM_EDGE_NEVER = 3  /* any other label (e.g. `string(&rune)` garbage,
                           bt_thomson.cpp:11): can never fire; kept for edge order */
};

#define RXM_MAX_CELLS 9u      /* README.md:22: cells are named 1..9              */
#define RXM_MAX_STATES 4096u  /* table format limit (kernels have their own)     */

/*
 * Structure-of-arrays transition table in CSR form.  States are numbered by the
 * ADDRESS RANK of the reference's Node* / MemoryNode* objects (that is the
 * iteration order of std::set<Node*> / std::set<MemoryState>, automata.h:12-13,
 * automata.cpp:122, mfa.cpp:206); a state's edges keep std::list order
 * (node.h:13,29).  Only states reachable from `start` (plus `finish`) appear.
 */
typedef struct rxm_tables {
    uint32_t abi_version;  /* RXM_ABI_VERSION */
    uint32_t kind;         /* RXM_KIND_NFA | RXM_KIND_MFA  (is_mfa of Regexp::compile) */
    uint32_t reversed;     /* Automata::is_reversed / MFA::is_reversed (automata.h:24,55) */
    uint32_t n_states;
    uint32_t n_edges;
    uint32_t start;
    uint32_t finish;
    uint32_t n_cells;      /* highest memory-cell id used by any edge, 0..9        */
    const uint32_t *edge_begin; /* [n_states + 1]  CSR row starts                 */
    const uint8_t *edge_kind;   /* [n_edges]  RXM_EDGE_*                          */
    const uint8_t *edge_sym;    /* [n_edges]  literal byte for RXM_EDGE_LIT, else 0 */
    const uint16_t *edge_to;    /* [n_edges]  target state                        */
    const uint16_t *edge_open;  /* [n_edges]  bit (k-1) set: memoryActions["k"] == open  (edge.h:29-32) */
    const uint16_t *edge_close; /* [n_edges]  bit (k-1) set: memoryActions["k"] == close */
} rxm_tables;

/* Text form of a table (one automaton), used for fixtures and for handing a
 * table from the flattening stage to another process.  rxm_tables_parse
 * allocates ONE block owned by the caller: release with rxm_tables_release.  */
int rxm_tables_format(const rxm_tables *t, char *buf, size_t buf_size, size_t *needed);
int rxm_tables_parse(const char *text, size_t len, rxm_tables **out);
void rxm_tables_release(rxm_tables *t);
int rxm_tables_validate(const rxm_tables *t);

/* ---- device side ---------------------------------------------------------- */
typedef struct rxm_matcher *rxm_handle;

/* Which kernel family the host planner picked for a table. */
enum {
    RXM_ENGINE_K1_DFA = 1,     /* memory-free automaton, determinised with the reference's exact step */
    RXM_ENGINE_K1_BITSET = 2,  /* memory-free automaton too large to determinise: the active set
                                  itself on the device (follow masks, or an edge walk)             */
    RXM_ENGINE_K2_THREAD = 3,  /* MFA, one thread per string, recursive walk (automata with more than 4 cells) */
    RXM_ENGINE_K3_WARP = 4,    /* MFA, one warp per string (large automata; takes over the strings that
                                  outgrow K4's per-thread sets)                                     */
    RXM_ENGINE_K4_THREAD = 5   /* MFA, one thread per string over host-compiled edge programs, repeated
                                  steps answered by their block compares alone                      */
};

typedef struct rxm_plan_info {
    uint32_t engine;        /* RXM_ENGINE_*                                        */
    uint32_t dfa_states;    /* K1_DFA: reachable active sets (incl. dead state)    */
    uint32_t dfa_classes;   /* K1_DFA: byte classes; K1_BITSET: classes of the follow masks,
                               0 when the edge-walking step is used                   */
    uint32_t exact_step_differs; /* memory-free: #(set,letter) pairs where the reference's
                                    `visited` step (automata.cpp:104-107) differs from the
                                    textbook step -- informational                   */
    uint32_t n_states;
    uint32_t n_edges;
    uint32_t n_cells;
    uint32_t reversed;
    uint32_t sm_count;
    uint32_t dfa_stride;    /* K1_DFA: input bytes per table lookup in the scan's interior: 1; 4 when all
                               literals lie in one 4-letter window and there are <= 64 sets; 8 when the
                               window has two letters.  Two-lookup tables (more than 128 sets; more than
                               64 with a two-letter window): 4 with a two-letter window while the stride
                               table fits shared memory (~4700 sets), else 1                           */
    uint32_t reserved[6];
} rxm_plan_info;

/* Copies `host_tables` (caller keeps ownership), plans, uploads to `device`. */
int rxm_tables_upload(const rxm_tables *host_tables, int device, rxm_handle *out);

/*
 * The same with the planner's choices overridden -- for tests and for comparing the engines; no reference
 * counterpart (the reference has one matcher per automaton kind).  `opts` may be NULL (= rxm_tables_upload).
 * An engine that cannot run the automaton is RXM_ERR_UNSUPPORTED, never a silent substitute.
 */
#define RXM_OPT_K1_NO_QUAD 1u   /* K1: one input byte per table lookup even where four would do          */
#define RXM_OPT_K1B_WALK 2u     /* K1 bit-set engine: the edge-walking step instead of the follow masks   */
#define RXM_OPT_INDEX_ORDER 4u  /* hand strings out by index, not in the tile sort's order                */
#define RXM_OPT_K1_NO_OCT 8u    /* K1: at most four input bytes per lookup even where eight would do      */
typedef struct rxm_upload_opts {
    uint32_t abi_version;  /* RXM_ABI_VERSION                                                             */
    uint32_t engine;       /* 0: the planner's choice; RXM_ENGINE_*: that engine                          */
    uint32_t flags;        /* RXM_OPT_*                                                                   */
    uint32_t k3_tile;      /* 0: the planner's choice; 8 / 16 / 32 lanes per string for K3                */
    uint32_t reserved[4];
} rxm_upload_opts;
int rxm_tables_upload_opts(const rxm_tables *host_tables, int device, const rxm_upload_opts *opts, rxm_handle *out);
int rxm_plan_query(rxm_handle h, rxm_plan_info *info);
int rxm_free(rxm_handle h);

/*
 * Match strings i = 0..n-1, string i = chars[offsets[i] .. offsets[i+1]).
 * Writes out_bits[i] = 0 or 1 (one BYTE per string, like the `0`/`1` lines of
 * match.cpp:29).  `chars`, `offsets`, `out_bits` may each be a device pointer
 * (on the handle's device) or a host pointer (pageable or pinned); host
 * buffers are staged through the handle's device workspace inside the call.
 * `stream` is a cudaStream_t (NULL = default stream).  With device pointers the
 * call is asynchronous on `stream` -- nothing is read back, every choice that
 * depends on the batch (string order, lanes per string) is made on the device;
 * with any host pointer it returns after the results are in `out_bits`.  Calls on
 * one handle must be serialised by the caller; different handles may be used
 * concurrently.  Workspace the handle owns (per-string records, lists) grows on
 * demand inside the call (cudaMalloc when a batch is larger than any before).
 *
 * Device buffers -- readable slack.  The kernels read whole ALIGNED words around a
 * string's bytes: the 32-byte-aligned blocks that hold its first and last byte
 * (K1, tokeniser) or the 8-byte ones (K3, K4).  So the device memory from
 * (chars + offsets[0]) rounded DOWN to 32 bytes up to (chars + offsets[n])
 * rounded UP to 32 bytes must be readable.  Any cudaMalloc'ed / pool-allocated
 * buffer satisfies this when `chars` is its start or 32-byte aligned inside it
 * (allocations are 256-byte granular); a sub-buffer that starts at an odd offset
 * must have its neighbours inside the same allocation.  Bytes outside the strings
 * never influence a result.
 */
int rxm_match_batch(rxm_handle h, const uint8_t *chars, const uint64_t *offsets, uint64_t n,
                    uint8_t *out_bits, void *stream);

/*
 * The same for RAW TEXT: `text[0 .. nbytes)` is split on the device into the tokens
 * `cin >> text` would deliver (matchers/match.cpp:22-23: runs of space, \t, \n, \v,
 * \f, \r separate tokens) and the tokens before the first token `exit`
 * (match.cpp:24; all of them if there is none) are matched in order:
 * out_bits[k] = 0/1 for token k, *n_tokens = how many there were, *saw_exit (may
 * be NULL) = 1 if the sentinel was met.  `text` and
 * `out_bits` are both host or both device pointers; `out_cap` is the room in
 * out_bits.  If there are more tokens than `out_cap` nothing is matched, *n_tokens is
 * set and the call returns RXM_ERR_INVALID -- call again with that much room.
 * nbytes must be below 4 GiB (cut larger inputs at whitespace).  The call synchronises
 * `stream` once (the token count comes back to the host); with device pointers the
 * matching itself is then asynchronous on `stream` as in rxm_match_batch.
 */
int rxm_match_text(rxm_handle h, const uint8_t *text, uint64_t nbytes, uint8_t *out_bits,
                   uint64_t out_cap, uint64_t *n_tokens, int *saw_exit, void *stream);

/*
 * Hint (optional; default 1): up to `handles` matchers are used at once on this device, each on
 * a stream of its own.  The MFA kernel K3 is persistent -- one launch occupies every block slot
 * of the device until its strings are handed out -- so a second handle's kernel would start
 * when the first one's blocks retire.  With the hint each K3 launch of `h` takes 1/handles of the
 * slots and the others run beside it (the other engines ignore it).  Worth it for batches
 * bounded by their longest string (the reference's attack strings: forward and -reverse
 * automaton of one expression side by side); throughput-bound batches are better left at 1.
 * No reference counterpart: the reference matches one expression per process
 * (matchers/match.cpp:10-32).
 */
int rxm_set_concurrency(rxm_handle h, uint32_t handles);

/* Number of kernels of this library launched through `h` so far. */
int rxm_launch_count(rxm_handle h, uint64_t *launches);

/* Strings whose simulation hit a kernel limit in the LAST rxm_match_batch on
 * `h` (frontier recursion depth).  Their out_bits are 0 and the call returns
 * RXM_ERR_OVERFLOW when the count is non-zero and the buffers were host
 * buffers; with device buffers query this after synchronising the stream.     */
int rxm_overflow_count(rxm_handle h, uint64_t *count);

const char *rxm_strerror(int status);
const char *rxm_last_cuda_error(void);

#ifdef __cplusplus
}
#endif
#endif /* RXM_H */
