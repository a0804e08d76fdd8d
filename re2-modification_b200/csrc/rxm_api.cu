// C-ABI launch layer (include/rxm.h) + kernel launches.  sm_100a only.
//
// No CPU fallback lives here: if the CUDA runtime cannot provide the requested
// device every compute entry point returns an error status.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/rxm.h"
#include "rxm_host_tables.hpp"
#include "rxm_kernels.cuh"
#include "rxm_mfa_dispatch.hpp"
#include "rxm_plan.hpp"

namespace {

thread_local char g_cuda_err[512] = "";

int cuda_fail(cudaError_t e, const char *what) {
    std::snprintf(g_cuda_err, sizeof g_cuda_err, "%s: %s", what, cudaGetErrorString(e));
    return RXM_ERR_CUDA;
}
#define CU(call)                                        \
    do {                                                \
        cudaError_t e_ = (call);                        \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
    } while (0)

template <class T>
int upload_vec(const std::vector<T> &v, T **dptr) {
    *dptr = nullptr;
    if (v.empty()) return RXM_OK;
    CU(cudaMalloc(reinterpret_cast<void **>(dptr), v.size() * sizeof(T)));
    CU(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return RXM_OK;
}

}  // namespace

struct rxm_matcher {
    int device = 0;
    int sm_count = 0;
    rxm::HostTables tables;
    rxm_plan_info info{};

    // K1: determinised automaton
    rxm::DfaPlan dfa;
    rxm::K1Tables k1{};
    uint8_t *d_k1_table = nullptr;   // direct: [256][SP] u8 ; classed: see rxm_kernels.cuh
    uint8_t *d_k1_accept = nullptr;
    uint32_t *d_k1b_edges = nullptr;  // K1B: packed edges (d_edge_begin holds the row starts)
    rxm::BitsetMasks bmasks;          // K1B: follow masks (bit-parallel step) when the table allows them
    uint64_t *d_k1b_ls = nullptr;
    uint8_t *d_k1b_class = nullptr;

    // K2: MFA tables
    uint16_t *d_edge_begin = nullptr;
    uint64_t *d_edges = nullptr;

    // K3: edge programs
    rxm::MfaProgram prog;
    rxm::ProgItem *d_items = nullptr;
    uint32_t *d_prog_begin = nullptr;
    uint32_t *d_prog_count = nullptr;
    uint32_t k3_tile = 32;  // lanes per string
    bool k3_tile_forced = false;
    // K4: the programs' item lists; strings that outgrow a thread's sets go to K3 through the redo list
    uint32_t *d_prog_lists = nullptr;  // K4's tables packed (rxm::k4_pack_lists)
    uint16_t *d_prog_sel = nullptr;
    uint32_t k4_maxl = 0;
    uint32_t *d_redo_list = nullptr;
    size_t cap_redo = 0;
    unsigned long long *d_redo_n = nullptr;
    uint32_t *d_pick = nullptr;  // [0] 1: the batch in hand is one of long strings (decided on the device)
    uint32_t sharing = 1;  // rxm_set_concurrency: handles that run at once on this device
    bool index_order = false;  // RXM_OPT_INDEX_ORDER

    // staging workspace for host buffers
    uint8_t *d_chars = nullptr;
    size_t cap_chars = 0;
    uint64_t *d_offsets = nullptr;
    uint8_t *d_bits = nullptr;
    size_t cap_n = 0;

    // K1 tile-sort workspace
    rxm::K1Rec *d_recs = nullptr;
    size_t cap_recs = 0;
    uint32_t *d_k1_counter = nullptr;    // scan kernel's task counter

    // tokeniser workspace (rxm_match_text)
    uint64_t *d_tok_begin = nullptr, *d_tok_end = nullptr;
    size_t cap_tok = 0;
    uint64_t *d_tok_masks = nullptr, *d_tok_counts = nullptr;
    size_t cap_tok_blocks = 0;
    unsigned long long *d_tok_result = nullptr;  // [2] token count, index of the first `exit`

    unsigned long long *d_overflow = nullptr;  // strings that hit a kernel limit
    uint64_t launches = 0;
    uint64_t last_overflow = 0;
};

static bool is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

extern "C" const char *rxm_last_cuda_error(void) { return g_cuda_err; }

extern "C" int rxm_tables_upload(const rxm_tables *host_tables, int device, rxm_handle *out) {
    return rxm_tables_upload_opts(host_tables, device, nullptr, out);
}

static int upload_impl(const rxm_tables *host_tables, int device, const rxm_upload_opts *opts, rxm_handle *out,
                       rxm_matcher **in_flight);

// never throws: an allocation failure inside the planner's containers is RXM_ERR_NOMEM, nothing leaks
extern "C" int rxm_tables_upload_opts(const rxm_tables *host_tables, int device, const rxm_upload_opts *opts,
                                      rxm_handle *out) {
    if (!out) return RXM_ERR_INVALID;
    *out = nullptr;
    rxm_matcher *in_flight = nullptr;
    try {
        return upload_impl(host_tables, device, opts, out, &in_flight);
    } catch (...) {
        if (in_flight) rxm_free(in_flight);
        *out = nullptr;
        return RXM_ERR_NOMEM;
    }
}

static int upload_impl(const rxm_tables *host_tables, int device, const rxm_upload_opts *opts, rxm_handle *out,
                       rxm_matcher **in_flight) {
    rxm_upload_opts o{};
    if (opts) {
        if (opts->abi_version != RXM_ABI_VERSION) return RXM_ERR_INVALID;
        o = *opts;
        if (o.engine > RXM_ENGINE_K4_THREAD || (o.k3_tile != 0 && o.k3_tile != 8 && o.k3_tile != 16 && o.k3_tile != 32))
            return RXM_ERR_INVALID;
    }
    int st = rxm_tables_validate(host_tables);
    if (st != RXM_OK) return st;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cuda_fail(e, "cudaGetDeviceCount");
        cudaGetLastError();
        return RXM_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= ndev) return RXM_ERR_NO_DEVICE;
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));

    rxm_matcher *m = new (std::nothrow) rxm_matcher();
    if (!m) return RXM_ERR_NOMEM;
    *in_flight = m;
    m->device = device;
    m->sm_count = prop.multiProcessorCount;
    m->tables.assign(*host_tables);
    const rxm_tables t = m->tables.view();
    m->info.n_states = t.n_states;
    m->info.n_edges = t.n_edges;
    m->info.n_cells = t.n_cells;
    m->info.reversed = t.reversed;
    m->info.sm_count = uint32_t(m->sm_count);
    std::string err;

    auto fail = [&](int code) {
        if (!err.empty()) std::snprintf(g_cuda_err, sizeof g_cuda_err, "%s", err.c_str());
        *in_flight = nullptr;
        rxm_free(m);
        return code;
    };

    m->index_order = (o.flags & RXM_OPT_INDEX_ORDER) != 0;
    if (o.engine != 0 &&
        (t.kind == RXM_KIND_NFA) != (o.engine == RXM_ENGINE_K1_DFA || o.engine == RXM_ENGINE_K1_BITSET)) {
        err = "the requested engine does not run this kind of automaton";
        return fail(RXM_ERR_UNSUPPORTED);
    }
    if (t.kind == RXM_KIND_NFA) {
        const bool force_bitset = o.engine == RXM_ENGINE_K1_BITSET;
        std::vector<uint8_t> table, accept;
        st = force_bitset ? RXM_ERR_UNSUPPORTED : rxm::plan_dfa(t, m->dfa, &err);
        // (a determinisation whose tables do not fit shared memory goes to the bit-set engine as well)
        if (st == RXM_OK) st = rxm::k1_build_tables(m->dfa, m->k1, table, accept, &err, (o.flags & RXM_OPT_K1_NO_QUAD) != 0,
                                                        (o.flags & RXM_OPT_K1_NO_OCT) != 0);
        if (st == RXM_ERR_UNSUPPORTED && o.engine != RXM_ENGINE_K1_DFA) {
            // too many active sets for a table (or forced): simulate the set itself (K1B)
            std::string berr;
            if (rxm::check_nfa_bitset(t, &berr) != RXM_OK) {
                if (err.empty() || force_bitset) err = berr;
                else err += "; bit-set engine: " + berr;
                return fail(RXM_ERR_UNSUPPORTED);
            }
            err.clear();
            std::vector<uint16_t> eb;
            std::vector<uint32_t> ed;
            rxm::k1b_build_tables(t, eb, ed);
            if ((st = upload_vec(eb, &m->d_edge_begin)) != RXM_OK) return fail(st);
            if ((st = upload_vec(ed, &m->d_k1b_edges)) != RXM_OK) return fail(st);
            rxm::plan_bitset_masks(t, m->bmasks);
            if (o.flags & RXM_OPT_K1B_WALK) m->bmasks.ok = false;  // the edge-walking step
            if (m->bmasks.ok && size_t(m->bmasks.n_classes) * t.n_states * 16 + 256 > 96 * 1024)
                m->bmasks.ok = false;  // the follow masks would not fit shared memory: decided here, not per batch
            if (m->bmasks.ok) {
                std::vector<uint8_t> cls(m->bmasks.byte_class, m->bmasks.byte_class + 256);
                if ((st = upload_vec(m->bmasks.ls, &m->d_k1b_ls)) != RXM_OK) return fail(st);
                if ((st = upload_vec(cls, &m->d_k1b_class)) != RXM_OK) return fail(st);
            }
            m->info.dfa_classes = m->bmasks.ok ? m->bmasks.n_classes : 0;  // 0: the edge-walking step
            m->info.engine = RXM_ENGINE_K1_BITSET;
            goto planned;
        }
        if (st != RXM_OK) return fail(st);
        if ((st = upload_vec(table, &m->d_k1_table)) != RXM_OK) return fail(st);
        if ((st = upload_vec(accept, &m->d_k1_accept)) != RXM_OK) return fail(st);
        m->info.engine = RXM_ENGINE_K1_DFA;
        m->info.dfa_states = m->dfa.n_states;
        m->info.dfa_classes = m->dfa.n_classes;
        m->info.dfa_stride = m->k1.quad == 2 ? 8 : (m->k1.quad ? 4 : 1);
        m->info.exact_step_differs = m->dfa.exact_step_differs;
    } else {
        st = rxm::check_mfa(t, &err);
        if (st != RXM_OK) return fail(st);
        if (t.n_states > RXM_MFA_MAX_STATES) {
            err = "MFA with more than " + std::to_string(RXM_MFA_MAX_STATES) + " states";
            return fail(RXM_ERR_UNSUPPORTED);
        }
        std::vector<uint16_t> eb(t.n_states + 1);
        for (uint32_t q = 0; q <= t.n_states; q++) eb[q] = uint16_t(t.edge_begin[q]);
        std::vector<uint64_t> er(t.n_edges);
        for (uint32_t i = 0; i < t.n_edges; i++)
            er[i] = rxm::pack_edge(t.edge_kind[i], t.edge_sym[i], t.edge_to[i], t.edge_open[i],
                                   t.edge_close[i]);
        if (t.n_edges > 8000) {
            err = "MFA with more than 8000 edges";
            return fail(RXM_ERR_UNSUPPORTED);
        }
        if ((st = upload_vec(eb, &m->d_edge_begin)) != RXM_OK) return fail(st);
        if ((st = upload_vec(er, &m->d_edges)) != RXM_OK) return fail(st);
        m->info.engine = RXM_ENGINE_K2_THREAD;
        // K3 (warp per string) needs the edge programs; automata they cannot express stay on K2
        std::string perr;
        if (o.engine != RXM_ENGINE_K2_THREAD && rxm::compile_programs(t, m->prog, &perr) == RXM_OK) {
            if ((st = upload_vec(m->prog.items, &m->d_items)) != RXM_OK) return fail(st);
            if ((st = upload_vec(m->prog.begin, &m->d_prog_begin)) != RXM_OK) return fail(st);
            if ((st = upload_vec(m->prog.count, &m->d_prog_count)) != RXM_OK) return fail(st);
            if ((st = upload_vec(m->prog.sel, &m->d_prog_sel)) != RXM_OK) return fail(st);
            {
                const std::vector<uint32_t> lists = rxm::k4_pack_lists(m->prog);  // K4's tables as one array
                if ((st = upload_vec(lists, &m->d_prog_lists)) != RXM_OK) return fail(st);
            }
            m->k4_maxl = rxm::k4_pool_for(t.n_states);  // slots per thread: the current set and the one being built share them
            // K4 (one thread per string) unless the automaton is large: its sets outgrow a thread's slots, K3 (one
            // warp per string, one slot per node) runs the whole batch
            const bool big = t.n_states > rxm::K4_MAX_STATES && o.engine != RXM_ENGINE_K4_THREAD;
            m->info.engine = (o.engine == RXM_ENGINE_K3_WARP || big) ? RXM_ENGINE_K3_WARP : RXM_ENGINE_K4_THREAD;
            // short programs: several strings per warp (a step's items fit one pass of the tile)
            m->k3_tile = m->prog.max_count <= 8 ? 8 : (m->prog.max_count <= 16 ? 16 : 32);
            if (o.k3_tile) {
                m->k3_tile = o.k3_tile;
                m->k3_tile_forced = true;
            }
        } else if (o.engine == RXM_ENGINE_K3_WARP || o.engine == RXM_ENGINE_K4_THREAD) {
            err = "the requested engine needs the edge programs, which cannot be built: " + perr;
            return fail(RXM_ERR_UNSUPPORTED);
        }
    }
planned:
    // [0] strings that hit a kernel limit, [1] K2/K3 work counter
    if (cudaMalloc(reinterpret_cast<void **>(&m->d_overflow), 2 * sizeof(unsigned long long)) != cudaSuccess)
        return fail(cuda_fail(cudaGetLastError(), "cudaMalloc overflow counter"));
    if (cudaMemset(m->d_overflow, 0, 2 * sizeof(unsigned long long)) != cudaSuccess)
        return fail(cuda_fail(cudaGetLastError(), "cudaMemset overflow counter"));
    *in_flight = nullptr;
    *out = m;
    return RXM_OK;
}

extern "C" int rxm_plan_query(rxm_handle h, rxm_plan_info *info) {
    if (!h || !info) return RXM_ERR_INVALID;
    *info = h->info;
    return RXM_OK;
}

extern "C" int rxm_free(rxm_handle h) {
    if (!h) return RXM_OK;
    cudaSetDevice(h->device);
    cudaFree(h->d_k1_table);
    cudaFree(h->d_k1_accept);
    cudaFree(h->d_k1b_edges);
    cudaFree(h->d_k1b_ls);
    cudaFree(h->d_k1b_class);
    cudaFree(h->d_edge_begin);
    cudaFree(h->d_edges);
    cudaFree(h->d_items);
    cudaFree(h->d_prog_begin);
    cudaFree(h->d_prog_count);
    cudaFree(h->d_prog_sel);
    cudaFree(h->d_prog_lists);
    cudaFree(h->d_redo_list);
    cudaFree(h->d_redo_n);
    cudaFree(h->d_pick);
    cudaFree(h->d_chars);
    cudaFree(h->d_offsets);
    cudaFree(h->d_bits);
    cudaFree(h->d_overflow);
    cudaFree(h->d_recs);
    cudaFree(h->d_k1_counter);
    cudaFree(h->d_tok_begin);
    cudaFree(h->d_tok_end);
    cudaFree(h->d_tok_masks);
    cudaFree(h->d_tok_counts);
    cudaFree(h->d_tok_result);
    delete h;
    return RXM_OK;
}

extern "C" int rxm_launch_count(rxm_handle h, uint64_t *launches) {
    if (!h || !launches) return RXM_ERR_INVALID;
    *launches = h->launches;
    return RXM_OK;
}

extern "C" int rxm_set_concurrency(rxm_handle h, uint32_t handles) {
    if (!h || handles == 0) return RXM_ERR_INVALID;
    h->sharing = handles;
    return RXM_OK;
}

extern "C" int rxm_overflow_count(rxm_handle h, uint64_t *count) {
    if (!h || !count) return RXM_ERR_INVALID;
    unsigned long long v = 0;
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpy(&v, h->d_overflow, sizeof v, cudaMemcpyDeviceToHost));
    h->last_overflow = v;
    *count = v;
    return RXM_OK;
}

// total_chars: offsets[n] - offsets[0] if the caller knows it, else ~0ull
// Tile-sorted order of a batch (long strings first, equal lengths together) for the engines that
// hand strings out one by one; *order stays null for small batches (and with RXM_K3_ORDER=index).
// long_strings: the batch is in the one-warp-per-string regime (mean length > 4096), where the
// longest strings bound the launch and 16 us of sorting pay at any batch size that outnumbers
// the strings in flight.
static int prepare_order(rxm_matcher *m, rxm::Spans spans, uint64_t n, cudaStream_t stream,
                         const rxm::K1Rec **order, int *launched, bool long_strings = false) {
    *order = nullptr;
    const uint64_t least = long_strings ? 512 : 4 * rxm::K1_TILE_STRINGS;
    if (n < least || n > 0xfffffff0ull || m->index_order) return RXM_OK;
    if (n > m->cap_recs) {
        cudaFree(m->d_recs);
        m->d_recs = nullptr;
        m->cap_recs = 0;
        const size_t want = size_t(n + (n >> 3) + 32);
        CU(cudaMalloc(reinterpret_cast<void **>(&m->d_recs), want * sizeof(rxm::K1Rec)));
        m->cap_recs = want;
    }
    if (!m->d_k1_counter) CU(cudaMalloc(reinterpret_cast<void **>(&m->d_k1_counter), 64 * sizeof(uint32_t)));
    const int st = rxm::k1_tilesort_launch(spans, n, m->d_recs, m->d_k1_counter, nullptr, m->sm_count, stream);
    if (st != RXM_OK) return st;
    *launched = 1;
    *order = m->d_recs;
    return RXM_OK;
}

static int launch_on_device(rxm_matcher *m, const uint8_t *d_chars, rxm::Spans spans,
                            uint64_t n, uint8_t *d_out, cudaStream_t stream, uint64_t total_chars) {
    CU(cudaMemsetAsync(m->d_overflow, 0, sizeof(unsigned long long), stream));
    if (n == 0) return RXM_OK;
    int launched = 0, launched_extra = 0;
    int st;
    if (m->info.engine == RXM_ENGINE_K1_DFA) {
        if (n > 0xfffffff0ull) return RXM_ERR_UNSUPPORTED;  // record index is 32 bits
        if (n > m->cap_recs) {
            cudaFree(m->d_recs);
            m->d_recs = nullptr;
            m->cap_recs = 0;
            const size_t want = size_t(n + (n >> 3) + 32);
            CU(cudaMalloc(reinterpret_cast<void **>(&m->d_recs), want * sizeof(rxm::K1Rec)));
            m->cap_recs = want;
        }
        if (!m->d_k1_counter) CU(cudaMalloc(reinterpret_cast<void **>(&m->d_k1_counter), 64 * sizeof(uint32_t)));
        rxm::K1Launch a{m->d_k1_table, m->d_k1_accept, d_chars, spans, n, d_out, m->d_recs,
                        m->d_k1_counter, m->d_overflow, m->sm_count, stream};
        st = rxm::k1_launch(m->k1, a, &launched);
    } else if (m->info.engine == RXM_ENGINE_K1_BITSET) {
        const rxm::K1Rec *order = nullptr;
        if ((st = prepare_order(m, spans, n, stream, &order, &launched_extra)) != RXM_OK) return st;
        if (m->bmasks.ok)
            st = rxm::k1b_mask_launch(m->d_k1b_ls, m->d_k1b_class, m->tables.n_states(), m->bmasks.n_classes,
                                      m->tables.start, m->bmasks.accept[0], m->bmasks.accept[1], m->tables.reversed,
                                      d_chars, spans, order, n, d_out, m->d_overflow, m->d_overflow + 1, m->sm_count,
                                      stream, &launched);
        else
        st = rxm::k1b_launch(m->d_edge_begin, m->d_k1b_edges, m->tables.n_states(), m->tables.n_edges(),
                             m->tables.start, m->tables.finish, m->tables.reversed, d_chars, spans, order, n, d_out,
                             m->d_overflow, m->d_overflow + 1, m->sm_count, stream, &launched);
    } else if (m->info.engine == RXM_ENGINE_K4_THREAD) {
        rxm::MfaView v{m->d_edge_begin, m->d_edges, m->tables.n_states(), m->tables.start,
                       m->tables.finish, m->tables.reversed};
        rxm::K4Prog kp{m->d_items, m->d_prog_lists, m->d_prog_sel, m->prog.n_cells, m->prog.n_classes,
                       uint32_t(m->prog.begin.size())};
        rxm::ProgView gp{m->d_items, m->d_prog_begin, m->d_prog_count, m->prog.n_cells};
        // Batches of long strings (mean length above 4096: few strings, each bounded by its own length) are
        // K3's, 32 lanes per string; everything else is K4's.  With host buffers the lengths are known here;
        // with device buffers the choice is made ON THE DEVICE (mfa_pick_kernel) and both kernels are launched
        // gated on it, so the call stays asynchronous on `stream`.
        const bool known = total_chars != ~0ull;
        const bool is_long = known && total_chars / n > rxm::kMfaLongMean;
        const uint32_t *gate = nullptr;
        if (!known) {
            if (!m->d_pick) CU(cudaMalloc(reinterpret_cast<void **>(&m->d_pick), 2 * sizeof(uint32_t)));
            if ((st = rxm::mfa_pick_launch(spans, n, m->d_pick, stream)) != RXM_OK) return st;
            launched_extra++;
            gate = m->d_pick;
        }
        const rxm::K1Rec *order = nullptr;
        if ((st = prepare_order(m, spans, n, stream, &order, &launched_extra, is_long || !known)) != RXM_OK) return st;
        if (!is_long) {
            // strings that outgrow a thread's slots (the current set and the one being built share them) are run by K3
            const bool redo = 2 * m->tables.n_states() > m->k4_maxl;
            if (redo) {
                if (n > 0xfffffff0ull) return RXM_ERR_UNSUPPORTED;
                if (n > m->cap_redo) {
                    cudaFree(m->d_redo_list);
                    m->d_redo_list = nullptr;
                    m->cap_redo = 0;
                    const size_t want = size_t(n + (n >> 3) + 32);
                    CU(cudaMalloc(reinterpret_cast<void **>(&m->d_redo_list), want * sizeof(uint32_t)));
                    m->cap_redo = want;
                }
                if (!m->d_redo_n) CU(cudaMalloc(reinterpret_cast<void **>(&m->d_redo_n), sizeof(unsigned long long)));
                CU(cudaMemsetAsync(m->d_redo_n, 0, sizeof(unsigned long long), stream));
            }
            st = rxm::k4_launch(v, kp, uint32_t(m->prog.items.size()), uint32_t(m->prog.begin.size()),
                                uint32_t(m->prog.sel.size()), m->tables.n_cells, m->k4_maxl, d_chars, spans, order, n, d_out,
                                m->d_overflow, m->d_overflow + 1, redo ? m->d_redo_list : nullptr, m->d_redo_n, m->sm_count,
                                m->sharing, stream, &launched, gate);
            if (st == RXM_OK && redo) {
                int l3 = 0;
                st = rxm::k3_launch(v, gp, uint32_t(m->prog.items.size()), uint32_t(m->prog.begin.size()), m->tables.n_cells,
                                    m->k3_tile, d_chars, spans, nullptr, n, d_out, m->d_overflow, m->d_overflow + 1,
                                    m->sm_count, m->sharing, stream, &l3, m->d_redo_list, m->d_redo_n);
                launched_extra += l3;
            }
        }
        if (st == RXM_OK && (is_long || !known)) {
            int l3 = 0;
            st = rxm::k3_launch(v, gp, uint32_t(m->prog.items.size()), uint32_t(m->prog.begin.size()), m->tables.n_cells,
                                32, d_chars, spans, order, n, d_out, m->d_overflow, m->d_overflow + 1, m->sm_count,
                                m->sharing, stream, &l3, nullptr, nullptr, gate, 1u);
            if (is_long) launched = l3;
            else launched_extra += l3;
        }
    } else if (m->info.engine == RXM_ENGINE_K3_WARP) {
        rxm::MfaView v{m->d_edge_begin, m->d_edges, m->tables.n_states(), m->tables.start,
                       m->tables.finish, m->tables.reversed};
        rxm::ProgView gp{m->d_items, m->d_prog_begin, m->d_prog_count, m->prog.n_cells};
        // Lanes per string: short programs let 2 or 4 strings share a warp, which pays for many short
        // strings; long strings want the whole warp (per-string latency, wide block compares).  With host
        // buffers the mean length decides here; with device buffers it is decided on the device and the two
        // tile widths are launched gated on it (no read-back: the call stays asynchronous on `stream`).
        uint32_t tile = m->k3_tile;
        const bool known = total_chars != ~0ull;
        const bool long_strings = known && total_chars / n > rxm::kMfaLongMean;
        const bool two_widths = !known && tile < 32 && !m->k3_tile_forced;
        if (long_strings && !m->k3_tile_forced) tile = 32;
        const uint32_t *gate = nullptr;
        if (two_widths) {
            if (!m->d_pick) CU(cudaMalloc(reinterpret_cast<void **>(&m->d_pick), 2 * sizeof(uint32_t)));
            if ((st = rxm::mfa_pick_launch(spans, n, m->d_pick, stream)) != RXM_OK) return st;
            launched_extra++;
            gate = m->d_pick;
        }
        const rxm::K1Rec *order = nullptr;
        if ((st = prepare_order(m, spans, n, stream, &order, &launched_extra, long_strings || !known)) != RXM_OK) return st;
        if (two_widths) {
            int l3 = 0;
            st = rxm::k3_launch(v, gp, uint32_t(m->prog.items.size()), uint32_t(m->prog.begin.size()), m->tables.n_cells,
                                32, d_chars, spans, order, n, d_out, m->d_overflow, m->d_overflow + 1, m->sm_count,
                                m->sharing, stream, &l3, nullptr, nullptr, gate, 1u);
            launched_extra += l3;
            if (st != RXM_OK) return st;
        }
        st = rxm::k3_launch(v, gp, uint32_t(m->prog.items.size()), uint32_t(m->prog.begin.size()),
                            m->tables.n_cells, tile, d_chars, spans, order, n, d_out, m->d_overflow, m->d_overflow + 1,
                            m->sm_count, m->sharing, stream, &launched, nullptr, nullptr, gate, 0u);
    } else {
        rxm::MfaView v{m->d_edge_begin, m->d_edges, m->tables.n_states(), m->tables.start,
                       m->tables.finish, m->tables.reversed};
        st = rxm::k2_launch(v, m->tables.n_cells, m->tables.n_edges(), d_chars, spans, n, d_out,
                            m->d_overflow, m->d_overflow + 1, m->sm_count, stream, &launched);
    }
    m->launches += uint64_t(launched + launched_extra);
    if (st != RXM_OK) return st;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
    return RXM_OK;
}

extern "C" int rxm_match_batch(rxm_handle h, const uint8_t *chars, const uint64_t *offsets,
                               uint64_t n, uint8_t *out_bits, void *stream_) {
    if (!h || !offsets || (n && !out_bits)) return RXM_ERR_INVALID;
    rxm_matcher *m = h;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CU(cudaSetDevice(m->device));

    const bool dev_off = is_device_ptr(offsets);
    const bool dev_out = is_device_ptr(out_bits);
    const bool dev_chars = chars ? is_device_ptr(chars) : dev_off;
    if (dev_off && dev_out && dev_chars)
        return launch_on_device(m, chars, rxm::csr_spans(offsets), n, out_bits, stream, ~0ull);
    if (dev_off || dev_out || (chars && dev_chars)) return RXM_ERR_INVALID;  // all host or all device

    // host buffers: stage through the handle's workspace
    const uint64_t total = offsets[n];
    if (offsets[0] != 0 && !chars) return RXM_ERR_INVALID;
    if (offsets[0] > total || (total && !chars)) return RXM_ERR_INVALID;
    if (total + 64 > m->cap_chars) {
        cudaFree(m->d_chars);
        m->d_chars = nullptr;
        m->cap_chars = 0;
        const size_t want = size_t(total + 64 + (total >> 3));
        CU(cudaMalloc(reinterpret_cast<void **>(&m->d_chars), want));
        m->cap_chars = want;
    }
    if (n + 1 > m->cap_n) {
        cudaFree(m->d_offsets);
        cudaFree(m->d_bits);
        m->d_offsets = nullptr;
        m->d_bits = nullptr;
        m->cap_n = 0;
        const size_t want = size_t(n + 1 + (n >> 3));
        CU(cudaMalloc(reinterpret_cast<void **>(&m->d_offsets), want * sizeof(uint64_t)));
        CU(cudaMalloc(reinterpret_cast<void **>(&m->d_bits), want));
        m->cap_n = want;
    }
    if (total) CU(cudaMemcpyAsync(m->d_chars, chars, total, cudaMemcpyHostToDevice, stream));
    CU(cudaMemcpyAsync(m->d_offsets, offsets, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));
    // the offsets are checked while the copy runs (1 M of them cost the host ~0.6 ms): no kernel may see a string
    // that starts behind its end or outside the copied bytes
    {
        bool ok = true;
        for (uint64_t i = 0; i < n; i++) ok &= offsets[i] <= offsets[i + 1];
        if (!ok) {
            cudaStreamSynchronize(stream);
            return RXM_ERR_INVALID;
        }
    }
    int st = launch_on_device(m, m->d_chars, rxm::csr_spans(m->d_offsets), n, m->d_bits, stream, total - offsets[0]);
    if (st != RXM_OK) return st;
    unsigned long long ovf = 0;
    if (n) CU(cudaMemcpyAsync(out_bits, m->d_bits, n, cudaMemcpyDeviceToHost, stream));
    CU(cudaMemcpyAsync(&ovf, m->d_overflow, sizeof ovf, cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    m->last_overflow = ovf;
    return ovf ? RXM_ERR_OVERFLOW : RXM_OK;
}

// Raw text in, bits out: the tokenisation of match.cpp:22-24 (`cin >> text` until "exit")
// runs on the device (rxm_tok.cu), its spans feed the same kernels as rxm_match_batch.
extern "C" int rxm_match_text(rxm_handle h, const uint8_t *text, uint64_t nbytes, uint8_t *out_bits,
                              uint64_t out_cap, uint64_t *n_tokens, int *saw_exit, void *stream_) {
    if (!h || !n_tokens || (nbytes && !text) || (out_cap && !out_bits)) return RXM_ERR_INVALID;
    *n_tokens = 0;
    if (saw_exit) *saw_exit = 0;
    rxm_matcher *m = h;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CU(cudaSetDevice(m->device));
    const bool dev_text = nbytes ? is_device_ptr(text) : is_device_ptr(out_bits);
    const bool dev_out = out_cap ? is_device_ptr(out_bits) : dev_text;
    if (dev_text != dev_out) return RXM_ERR_INVALID;  // all host or all device
    if (nbytes >= (1ull << 32) - 64) return RXM_ERR_UNSUPPORTED;  // feed larger inputs in pieces cut at whitespace

    const uint8_t *d_text = text;
    if (!dev_text) {  // stage the text (and later the bits) through the handle's workspace
        if (nbytes + 64 > m->cap_chars) {
            cudaFree(m->d_chars);
            m->d_chars = nullptr;
            m->cap_chars = 0;
            const size_t want = size_t(nbytes + 64 + (nbytes >> 3));
            CU(cudaMalloc(reinterpret_cast<void **>(&m->d_chars), want));
            m->cap_chars = want;
        }
        if (nbytes) CU(cudaMemcpyAsync(m->d_chars, text, nbytes, cudaMemcpyHostToDevice, stream));
        d_text = m->d_chars;
    }
    if (!d_text) d_text = reinterpret_cast<const uint8_t *>(m->d_overflow);  // empty text: any valid address
    const uint64_t cap = out_cap ? out_cap : 1;
    if (cap > m->cap_tok) {
        cudaFree(m->d_tok_begin);
        cudaFree(m->d_tok_end);
        m->d_tok_begin = m->d_tok_end = nullptr;
        m->cap_tok = 0;
        const size_t want = size_t(cap + (cap >> 3) + 16);
        CU(cudaMalloc(reinterpret_cast<void **>(&m->d_tok_begin), want * sizeof(uint64_t)));
        CU(cudaMalloc(reinterpret_cast<void **>(&m->d_tok_end), want * sizeof(uint64_t)));
        m->cap_tok = want;
    }
    const uint64_t blocks = rxm::tok_blocks(nbytes);
    if (blocks > m->cap_tok_blocks) {
        cudaFree(m->d_tok_masks);
        cudaFree(m->d_tok_counts);
        m->d_tok_masks = m->d_tok_counts = nullptr;
        m->cap_tok_blocks = 0;
        const size_t want = size_t(blocks + (blocks >> 2) + 16);
        CU(cudaMalloc(reinterpret_cast<void **>(&m->d_tok_masks), want * 256 * sizeof(uint64_t)));
        CU(cudaMalloc(reinterpret_cast<void **>(&m->d_tok_counts), want * 9 * sizeof(uint64_t)));
        m->cap_tok_blocks = want;
    }
    if (!m->d_tok_result) CU(cudaMalloc(reinterpret_cast<void **>(&m->d_tok_result), 64));
    rxm::TokWork w{m->d_tok_masks, m->d_tok_counts, m->cap_tok_blocks, m->d_tok_result};
    int launched = 0;
    // one span more than the caller has room for: an `exit` standing right behind the last token
    // that fits must still be seen as the sentinel
    int st = rxm::tok_launch(d_text, nbytes, m->d_tok_begin, m->d_tok_end, out_cap + 1, w, m->sm_count, stream, &launched);
    m->launches += uint64_t(launched);
    if (st != RXM_OK) return st;
    unsigned long long res[2] = {0, 0};
    CU(cudaMemcpyAsync(res, w.d_result, sizeof res, cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    const uint64_t n = res[1] < res[0] ? res[1] : res[0];  // tokens before the first "exit"
    *n_tokens = n;
    if (saw_exit) *saw_exit = res[1] < res[0] ? 1 : 0;
    if (n > out_cap) return RXM_ERR_INVALID;  // *n_tokens says how many bits the caller must provide
    if (n == 0) return RXM_OK;

    const rxm::Spans spans{m->d_tok_begin, m->d_tok_end};
    if (dev_text) return launch_on_device(m, d_text, spans, n, out_bits, stream, nbytes);
    if (n > m->cap_n) {
        cudaFree(m->d_offsets);
        cudaFree(m->d_bits);
        m->d_offsets = nullptr;
        m->d_bits = nullptr;
        m->cap_n = 0;
        const size_t want = size_t(n + 1 + (n >> 3));
        CU(cudaMalloc(reinterpret_cast<void **>(&m->d_offsets), want * sizeof(uint64_t)));
        CU(cudaMalloc(reinterpret_cast<void **>(&m->d_bits), want));
        m->cap_n = want;
    }
    st = launch_on_device(m, d_text, spans, n, m->d_bits, stream, nbytes);
    if (st != RXM_OK) return st;
    unsigned long long ovf = 0;
    CU(cudaMemcpyAsync(out_bits, m->d_bits, n, cudaMemcpyDeviceToHost, stream));
    CU(cudaMemcpyAsync(&ovf, m->d_overflow, sizeof ovf, cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    m->last_overflow = ovf;
    return ovf ? RXM_ERR_OVERFLOW : RXM_OK;
}
