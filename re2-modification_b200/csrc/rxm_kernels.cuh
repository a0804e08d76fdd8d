// Kernel-side declarations shared by rxm_api.cu and rxm_kernels.cu.
#ifndef RXM_KERNELS_CUH
#define RXM_KERNELS_CUH

#ifndef RXM_SIMT_HOST
#include <cuda_runtime.h>
// kernel<<<grid, block, dynamic shared memory, stream>>>(...) and the kernel's view of that memory.
// tests/hostsim/simt_shim.hpp defines both for the host, so that the kernel sources AND their launch
// functions run in the CPU test tier on a SIMT emulator.
#define RXM_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define RXM_DYN_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#define RXM_DYN_SMEM_128(name) extern __shared__ __align__(128) uint8_t name[]
#endif

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/rxm.h"
#include "rxm_mfa_core.cuh"
#include "rxm_k4_core.cuh"
#include "rxm_plan.hpp"

namespace rxm {

// Where the strings of a batch are: string i = chars[begin[i] .. end[i]).  A CSR offsets array
// is {offsets, offsets + 1}; the device tokeniser (rxm_tok.cu) fills two separate arrays.
struct Spans {
    const uint64_t *begin;
    const uint64_t *end;
};
inline Spans csr_spans(const uint64_t *offsets) { return Spans{offsets, offsets + 1}; }

// ---- K1: determinised memory-free automaton ------------------------------------------
enum : uint32_t {
    K1_DIRECT = 0,   // T[byte][SP] u8, SP = 2^log2sp >= n_states, one shared-memory lookup per byte
    K1_CLASSED = 1   // cmap[256] u8 then trans[class][n_states] u16: two lookups per byte
};

struct K1Tables {
    uint32_t mode;
    uint32_t log2sp;       // K1_DIRECT
    uint32_t n_states;
    uint32_t n_classes;
    uint32_t start;
    uint32_t reversed;
    uint32_t table_bytes;  // bytes of the device table blob (copied to shared memory)
    uint32_t accept_bytes;
    uint32_t quad;         // K1_DIRECT with SP <= 64: 1 = the quad table Q[SP][256] (four bytes per lookup, 4-letter
                           // window) follows T in the blob, 2 = the oct table O[SP][256] (eight bytes, 2-letter window).
                           // K1_CLASSED with a 2-letter window: 1 = Q[n_states][16] u16 (four bytes per lookup, one bit
                           // per letter) at multi_off in the blob
    uint32_t quad_lo;      // lowest byte value of the window
    uint32_t multi_off;    // K1_CLASSED: byte offset of the stride table in the blob (16-byte aligned)
};

int k1_build_tables(const DfaPlan &p, K1Tables &kt, std::vector<uint8_t> &table,
                    std::vector<uint8_t> &accept, std::string *err, bool no_quad = false,
                    bool no_oct = false);

// One record per string, written by the tile sort: descending length bucket within each tile.
struct __align__(16) K1Rec {
    unsigned long long start;  // byte offset into chars
    uint32_t len;
    uint32_t idx;              // original string index (where the result bit goes)
};
constexpr uint32_t K1_BUCKETS = 2048;
constexpr uint32_t K1_TILE_STRINGS = 4096;  // strings per tile of the sort

struct K1Launch {
    const uint8_t *d_table;
    const uint8_t *d_accept;
    const uint8_t *d_chars;
    Spans spans;
    uint64_t n;
    uint8_t *d_out;
    K1Rec *d_recs;               // [n]        workspace
    uint32_t *d_task_counter;    // [1]        workspace
    unsigned long long *d_overflow;
    int sm_count;
    cudaStream_t stream;
};

int k1_launch(const K1Tables &kt, const K1Launch &a, int *launched);
// The tile sort alone (K3 uses its order to hand out the long strings first); d_overflow may be null.
int k1_tilesort_launch(Spans spans, uint64_t n, K1Rec *d_recs, uint32_t *d_counter, unsigned long long *d_overflow,
                       int sm_count, cudaStream_t stream);

// ---- K1B: memory-free automaton as a bit set (fallback when the determinisation is too large) ----
void k1b_build_tables(const rxm_tables &t, std::vector<uint16_t> &edge_begin, std::vector<uint32_t> &edges);
int k1b_launch(const uint16_t *d_eb, const uint32_t *d_ed, uint32_t n_states, uint32_t n_edges, uint32_t start,
               uint32_t finish, uint32_t reversed, const uint8_t *d_chars, Spans spans,
               const K1Rec *d_recs /* tile-sorted order, or null */, uint64_t n, uint8_t *d_out,
               unsigned long long *d_overflow, unsigned long long *d_next, int sm_count, cudaStream_t stream,
               int *launched);

int k1b_mask_launch(const uint64_t *d_ls, const uint8_t *d_class, uint32_t n_states, uint32_t n_classes, uint32_t start,
                    uint64_t acc_lo, uint64_t acc_hi, uint32_t reversed, const uint8_t *d_chars, Spans spans,
                    const K1Rec *d_recs, uint64_t n, uint8_t *d_out, unsigned long long *d_overflow,
                    unsigned long long *d_next, int sm_count, cudaStream_t stream, int *launched);

// ---- K2: MFA, one thread per string ------------------------------------------------------
int k2_launch(const MfaView &dev_view, uint32_t n_cells, uint32_t n_edges, const uint8_t *d_chars,
              Spans spans, uint64_t n, uint8_t *d_out, unsigned long long *d_overflow,
              unsigned long long *d_next /* work counter */, int sm_count, cudaStream_t stream,
              int *launched);

// ---- K3: MFA, one warp per string, over host-compiled edge programs --------------------------
int k3_launch(const MfaView &v, const ProgView &gp, uint32_t n_items, uint32_t n_keys, uint32_t n_cells,
              uint32_t tile /* lanes per string: 8, 16 or 32 */, const uint8_t *d_chars, Spans spans,
              const K1Rec *d_recs /* tile-sorted order, or null: index order */, uint64_t n, uint8_t *d_out,
              unsigned long long *d_overflow, unsigned long long *d_next, int sm_count,
              uint32_t sharing /* handles running at once on the device: the grid takes 1/sharing of the block slots */,
              cudaStream_t stream, int *launched,
              const uint32_t *d_list = nullptr /* run strings d_list[0 .. *d_list_n) instead of 0 .. n (n = the list's capacity) */,
              const unsigned long long *d_list_n = nullptr,
              const uint32_t *d_gate = nullptr /* run only if (*d_gate != 0) == (gate_want != 0): mfa_pick_launch */,
              uint32_t gate_want = 0);

// Batches of long strings (mean length above kMfaLongMean) go to K3 with 32 lanes per string, the others to
// K4.  When only the device knows the lengths the choice is made there: d_flag[0] <- 1 (long) / 0, and both
// kernels are launched gated on it -- no read-back, the call stays asynchronous.
constexpr uint64_t kMfaLongMean = 4096;
int mfa_pick_launch(Spans spans, uint64_t n, uint32_t *d_flag, cudaStream_t stream);

// ---- K4: MFA, one thread per string, over the same edge programs (rxm_k4_core.cuh) ------------
// maxl = live configurations a thread keeps per set; a string that needs more is appended to
// d_redo_list (capacity n) / *d_redo_n for K3 to run, or counted in d_overflow when the list is null.
int k4_launch(const MfaView &v, const K4Prog &gp, uint32_t n_items, uint32_t n_keys, uint32_t n_sel, uint32_t n_cells,
              uint32_t maxl, const uint8_t *d_chars, Spans spans, const K1Rec *d_recs, uint64_t n, uint8_t *d_out,
              unsigned long long *d_overflow, unsigned long long *d_next, uint32_t *d_redo_list,
              unsigned long long *d_redo_n, int sm_count, uint32_t sharing, cudaStream_t stream, int *launched,
              const uint32_t *d_gate = nullptr /* skip the launch's work if *d_gate != 0 (a batch of long strings) */);

// ---- tokeniser: whitespace-delimited text -> spans (rxm_tok.cu) ------------------------------
// Same token boundaries as `cin >> text` (matchers/match.cpp:22-23): whitespace is
// space, \t, \n, \v, \f, \r.  Writes begin/end of the first `cap` tokens and, into
// d_result[0..1], the number of tokens found and the index of the first token equal to
// "exit" (the sentinel of match.cpp:24; the token count if there is none).
struct TokWork {
    uint64_t *d_masks;       // [blocks * 256] one whitespace bit per input position
    uint64_t *d_counts;      // [blocks_cap] per 16 KB block (starts << 32 | ends), then their exclusive prefix;
                             // followed by [blocks_cap * 8]: each 2 KB warp piece's offset inside its block
    uint64_t blocks_cap;     // capacity of both, in blocks
    unsigned long long *d_result;  // [2]
};
uint64_t tok_blocks(uint64_t nbytes);  // 16 KB pieces a text of nbytes needs
int tok_launch(const uint8_t *d_text, uint64_t nbytes, uint64_t *d_begin, uint64_t *d_end, uint64_t cap,
               const TokWork &w, int sm_count, cudaStream_t stream, int *launched);

}  // namespace rxm
#endif
