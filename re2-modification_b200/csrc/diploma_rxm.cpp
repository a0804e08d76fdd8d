// diploma_rxm -- the reference's `-match` entry point with the simulation on the GPU.
//
// Mirrors main.cpp:9-45 + matchers/match.cpp:10-32 of the reference: same flags,
// same banners on stdout (they come from the reference's own Regexp::compile),
// same protocol (first token = expression, then one `0`/`1` line per
// whitespace-delimited token until the token `exit`).  What changes is WHO
// simulates: tokens are collected into batches and matched by librxm
// (include/rxm.h) on the device; MFA::match / Automata::match are never called
// and there is no CPU fallback -- if the device path fails the program exits
// non-zero.
//
//   diploma_rxm -match [-all|-bnf|-reverse|-ssnf|-log ...]          stdin protocol
//   diploma_rxm -match [flags] -batch IN.rxmb OUT.bits [-regex R]   file batch route
//       IN:  "RXMBATCH" | u64 n | u64 total | u64 offsets[n+1] | u8 chars[total]
//       OUT: n bytes, 0/1
//   extra: -device D (default 0), -chunk N (tokens per device batch, default 1<<20)
//
// Differences from the reference that are deliberate: results are printed when a
// batch is flushed (at `exit`, EOF, or every -chunk tokens) instead of after each
// token; EOF ends the loop (the reference spins forever, match.cpp:23-31).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "regex/regex.h"  // reference header (-I<reference root>)

#include "rxm_flatten.hpp"

static int flush_batch(rxm_handle h, std::vector<uint8_t> &chars, std::vector<uint64_t> &off) {
    const uint64_t n = off.size() - 1;
    if (n == 0) return RXM_OK;
    std::vector<uint8_t> bits(n);
    const int st = rxm_match_batch(h, chars.data(), off.data(), n, bits.data(), nullptr);
    if (st != RXM_OK) return st;
    for (uint64_t i = 0; i < n; i++) std::cout << int(bits[i]) << "\n";  // match.cpp:29
    std::cout.flush();
    chars.clear();
    off.assign(1, 0);
    return RXM_OK;
}

int main(int argc, char **argv) {
    if (argc < 2 || std::strcmp(argv[1], "-match") != 0) {
        std::fprintf(stderr, "usage: diploma_rxm -match [-all|-bnf|-reverse|-ssnf] "
                             "[-batch IN OUT] [-regex R] [-device D] [-chunk N]\n");
        return 2;
    }
    bool bnf = false, reverse = false, ssnf = false;
    const char *batch_in = nullptr, *batch_out = nullptr;
    std::string regex;
    bool have_regex = false;
    int device = 0;
    uint64_t chunk = uint64_t(1) << 20;
    // main.cpp:19-40
    if (argc > 2 && std::strcmp(argv[2], "-all") == 0) bnf = reverse = ssnf = true;
    for (int i = 2; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "-bnf") bnf = true;
        else if (a == "-reverse") { reverse = true; bnf = true; }
        else if (a == "-ssnf") ssnf = true;
        else if (a == "-regex" && i + 1 < argc) { regex = argv[++i]; have_regex = true; }
        else if (a == "-batch" && i + 2 < argc) { batch_in = argv[++i]; batch_out = argv[++i]; }
        else if (a == "-device" && i + 1 < argc) device = std::atoi(argv[++i]);
        else if (a == "-chunk" && i + 1 < argc) chunk = std::strtoull(argv[++i], nullptr, 10);
    }
    if (!have_regex && !(std::cin >> regex)) return 2;  // main.cpp:42-43

    Regexp *re = Regexp::parse_regexp(regex);                        // match.cpp:12
    bool is_mfa = false;
    Automata *automata = re->compile(is_mfa, reverse, bnf, ssnf);    // match.cpp:15 (banners, .dot files)
    std::cout.flush();

    rxm::HostTables host;
    std::string err;
    int st = rxm::flatten(automata, is_mfa, host, &err);
    if (st != RXM_OK) {
        std::fprintf(stderr, "diploma_rxm: flatten: %s (%s)\n", rxm_strerror(st), err.c_str());
        return 3;
    }
    const rxm_tables t = host.view();
    rxm_handle h = nullptr;
    st = rxm_tables_upload(&t, device, &h);
    if (st != RXM_OK) {
        std::fprintf(stderr, "diploma_rxm: upload: %s (%s)\n", rxm_strerror(st), rxm_last_cuda_error());
        return 3;
    }

    if (batch_in) {
        FILE *f = std::fopen(batch_in, "rb");
        char magic[8];
        uint64_t n = 0, total = 0;
        if (!f || std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "RXMBATCH", 8) != 0 ||
            std::fread(&n, 8, 1, f) != 1 || std::fread(&total, 8, 1, f) != 1) {
            std::fprintf(stderr, "diploma_rxm: cannot read %s\n", batch_in);
            return 2;
        }
        std::vector<uint64_t> off(n + 1);
        std::vector<uint8_t> chars(total), bits(n);
        if (std::fread(off.data(), 8, n + 1, f) != n + 1 ||
            (total && std::fread(chars.data(), 1, total, f) != total)) {
            std::fprintf(stderr, "diploma_rxm: short read on %s\n", batch_in);
            return 2;
        }
        std::fclose(f);
        st = rxm_match_batch(h, chars.data(), off.data(), n, bits.data(), nullptr);
        if (st != RXM_OK) {
            std::fprintf(stderr, "diploma_rxm: match: %s (%s)\n", rxm_strerror(st), rxm_last_cuda_error());
            return 3;
        }
        FILE *o = std::fopen(batch_out, "wb");
        if (!o) return 2;
        std::fwrite(bits.data(), 1, n, o);
        std::fclose(o);
        rxm_free(h);
        return 0;
    }

    std::vector<uint8_t> chars;
    std::vector<uint64_t> off(1, 0);
    std::string text;
    while (std::cin >> text) {  // match.cpp:22-31
        if (text == "exit") break;
        chars.insert(chars.end(), text.begin(), text.end());
        off.push_back(chars.size());
        if (off.size() - 1 >= chunk && (st = flush_batch(h, chars, off)) != RXM_OK) break;
    }
    if (st == RXM_OK) st = flush_batch(h, chars, off);
    if (st != RXM_OK) {
        std::fprintf(stderr, "diploma_rxm: match: %s (%s)\n", rxm_strerror(st), rxm_last_cuda_error());
        return 3;
    }
    rxm_free(h);
    return 0;
}
