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
//   diploma_rxm -match N [-maxlen L]                                 benchmark route (example_runner.cpp)
//   extra: -device D (default 0), -chunk N (bytes of input per device batch, default 256 MiB)
//
// The tokens are found ON THE DEVICE (rxm_match_text): stdin is read as raw bytes and
// shipped as it is.  Differences from the reference that are deliberate: results are
// printed when a piece is flushed (at `exit`, end of input, every -chunk bytes, or per
// line on a terminal) instead of after each token; end of input ends the loop (the
// reference spins forever, match.cpp:23-31).
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <string>
#include <vector>

// read(2) / isatty(3) under private names: <unistd.h> cannot be included next to the reference's
// edge.h, whose global `enum MemoryAction { open, close }` (edge.h:29-32) collides with close(2).
extern "C" long rxm_posix_read(int, void *, unsigned long) __asm__("read");
extern "C" int rxm_posix_isatty(int) __asm__("isatty");

#include "regex/regex.h"  // reference header (-I<reference root>)

#include "rxm_flatten.hpp"

// Matches the tokens of text[0..len) on the device (tokenised there too) and prints one
// `0`/`1` line per token (match.cpp:29).  *done <- the `exit` sentinel was met (match.cpp:24).
static int match_piece(rxm_handle h, const uint8_t *text, size_t len, bool *done) {
    std::vector<uint8_t> bits((len + 1) / 2 + 1);
    uint64_t n = 0;
    int saw_exit = 0;
    const int st = rxm_match_text(h, text, len, bits.data(), bits.size(), &n, &saw_exit, nullptr);
    if (st != RXM_OK) return st;
    std::string outbuf;
    outbuf.reserve(size_t(n) * 2);
    for (uint64_t i = 0; i < n; i++) {
        outbuf.push_back(bits[i] ? '1' : '0');
        outbuf.push_back('\n');
    }
    std::cout << outbuf;
    std::cout.flush();
    *done = saw_exit != 0;
    return RXM_OK;
}

// ---- `-match N`: the reference's benchmark route (matchers/example_runner.cpp:84-151) -------------
// Same inputs (test/example_N/regexp.txt, pump.txt relative to the working directory), same three
// automata (plain / bnf / reverse, all with -ssnf: example_runner.cpp:109-111), same CUMULATIVE
// attack strings (:123 appends to `prefix`), same stop rule (a measurement >= 0.5 s ends that
// automaton's series, :76-81) and the same `len seconds` lines in diploma_results.txt,
// diploma_bnf_results.txt and diploma_reverse_results.txt -- but each string is matched by the
// device kernels (one rxm_match_batch of one string, host buffers, wall clock around the call).
static std::string pumped_string(int n, const std::vector<std::string> &pump_v) {  // example_runner.cpp:15-29
    const int pump_count = int(pump_v.size()) / 2 + 1;
    const int del_count = int(pump_v.size()) - pump_count;
    const int tmp_len = int(pump_v[0].length());
    std::string res = pump_v[0];
    while (int(res.length()) + tmp_len < (n - del_count) / pump_count) res += pump_v[0];
    std::string out;
    for (int i = 0; i < del_count; i++) out += res + pump_v[1];
    out += res;
    return out;
}

static std::vector<std::string> split_commas(const std::string &str) {  // example_runner.cpp:31-44
    std::vector<std::string> parts;
    size_t start = 0;
    for (size_t i = 0; i <= str.size(); i++)
        if (i == str.size() || str[i] == ',') {
            parts.push_back(str.substr(start, i - start));
            start = i + 1;
        }
    return parts;
}

static bool read_line(FILE *f, std::string &out) {
    out.clear();
    int c;
    bool any = false;
    while ((c = std::fgetc(f)) != EOF) {
        any = true;
        if (c == '\n') break;
        out.push_back(char(c));
    }
    return any;
}

// bits_path (additive, `-bits FILE`): also write, per cumulative string, `RUNNER <len> <bit plain> <bit bnf>
// <bit reverse>` and -- for the same string before its failing suffix is appended -- a `RUNNER_NS` line:
// the format of oracle/ref_runner.cpp, which prints the reference's own bits for the same loop.
static int run_configuration_examples_gpu(const std::string &number, int device, size_t max_len, const char *bits_path) {
    const std::string dir = "test/example_" + number + "/";
    FILE *regex_file = std::fopen((dir + "regexp.txt").c_str(), "r");
    FILE *pump_file = std::fopen((dir + "pump.txt").c_str(), "r");
    if (!regex_file || !pump_file) {
        std::fprintf(stderr, "diploma_rxm: cannot open %sregexp.txt / pump.txt\n", dir.c_str());
        return 2;
    }
    std::string pump_s, suffix, prefix, regexp_str;
    read_line(pump_file, pump_s);
    read_line(pump_file, suffix);
    read_line(pump_file, prefix);
    read_line(regex_file, regexp_str);
    std::fclose(regex_file);
    std::fclose(pump_file);
    const std::vector<std::string> pump = split_commas(pump_s);
    std::cout << regexp_str << std::endl;  // :105

    std::string parse_copy = regexp_str;   // parse_regexp consumes its argument (parser.cpp:68)
    Regexp *regexp = Regexp::parse_regexp(parse_copy);
    regexp->is_backref_correct();
    const bool flags[3][3] = {{false, false, true}, {false, true, true}, {true, true, true}};  // reverse, bnf, ssnf (:109-111)
    const char *names[3] = {"diploma_results.txt", "diploma_bnf_results.txt", "diploma_reverse_results.txt"};
    rxm_handle h[3] = {nullptr, nullptr, nullptr};
    FILE *out[3] = {nullptr, nullptr, nullptr};
    for (int a = 0; a < 3; a++) {
        bool is_mfa = true;
        Automata *automata = regexp->compile(is_mfa, flags[a][0], flags[a][1], flags[a][2]);
        std::cout.flush();
        rxm::HostTables host;
        std::string err;
        int st = rxm::flatten(automata, is_mfa, host, &err);
        if (st != RXM_OK) {
            std::fprintf(stderr, "diploma_rxm: flatten: %s (%s)\n", rxm_strerror(st), err.c_str());
            return 3;
        }
        const rxm_tables t = host.view();
        st = rxm_tables_upload(&t, device, &h[a]);
        if (st != RXM_OK) {
            std::fprintf(stderr, "diploma_rxm: upload: %s (%s)\n", rxm_strerror(st), rxm_last_cuda_error());
            return 3;
        }
        out[a] = std::fopen((dir + names[a]).c_str(), "w");
        if (!out[a]) return 2;
    }
    {   // first use pays for context and workspace set-up: not part of any measurement
        const uint8_t w[1] = {'a'};
        const uint64_t off[2] = {0, 1};
        uint8_t bit;
        for (int a = 0; a < 3; a++) rxm_match_batch(h[a], w, off, 1, &bit, nullptr);
    }
    FILE *bits_file = bits_path ? std::fopen(bits_path, "w") : nullptr;
    if (bits_path && !bits_file) return 2;
    auto bits_line = [&](const char *tag, const std::string &str) {
        std::fprintf(bits_file, "%s %zu", tag, str.length());
        for (int a = 0; a < 3; a++) {
            const uint64_t off[2] = {0, uint64_t(str.length())};
            uint8_t bit = 0;
            const int st = rxm_match_batch(h[a], reinterpret_cast<const uint8_t *>(str.data()), off, 1, &bit, nullptr);
            if (st == RXM_OK) std::fprintf(bits_file, " %d", int(bit));
            else std::fprintf(bits_file, " E%d", st);
        }
        std::fprintf(bits_file, "\n");
    };
    bool timeouted[3] = {false, false, false};
    int count = 0;
    long long pump_size = 500;
    size_t len = prefix.length() + size_t(pump_size) + suffix.length();
    int rc = 0;
    while ((!timeouted[0] || !timeouted[1] || !timeouted[2]) && len < max_len) {  // :120
        if (pump_size > (1ll << 30)) break;
        const std::string no_suffix = bits_file ? prefix + pumped_string(int(pump_size), pump) : std::string();
        prefix.append(pumped_string(int(pump_size), pump)).append(suffix);  // :123 -- cumulative
        const std::string &input_str = prefix;
        len = input_str.length();
        if (bits_file && len <= max_len) {
            bits_line("RUNNER", input_str);
            bits_line("RUNNER_NS", no_suffix);
        }
        pump_size += pump_size;
        for (int a = 0; a < 3; a++) {
            if (timeouted[a]) continue;
            const uint64_t off[2] = {0, uint64_t(len)};
            uint8_t bit = 0;
            const auto t0 = std::chrono::steady_clock::now();
            const int st = rxm_match_batch(h[a], reinterpret_cast<const uint8_t *>(input_str.data()), off, 1, &bit, nullptr);
            const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (st != RXM_OK) {
                std::fprintf(stderr, "diploma_rxm: %s, length %zu: %s (%s)\n", names[a], len, rxm_strerror(st),
                             rxm_last_cuda_error());
                timeouted[a] = true;
                if (st != RXM_ERR_OVERFLOW) rc = 3;
                continue;
            }
            if (seconds >= 0.5) timeouted[a] = true;                                  // :78-79
            if (seconds < 1) std::fprintf(out[a], "%zu %g\n", len, seconds);           // :80-81
            if (a == 2) count++;                                                       // :140
        }
        if (count % 10 == 0) pump_size *= 2;  // :143-144
    }
    for (int a = 0; a < 3; a++) {
        std::fclose(out[a]);
        rxm_free(h[a]);
    }
    if (bits_file) std::fclose(bits_file);
    return rc;
}

static inline bool is_ws(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }  // what `cin >>` skips

int main(int argc, char **argv) {
    if (argc < 2 || std::strcmp(argv[1], "-match") != 0) {
        std::fprintf(stderr, "usage: diploma_rxm -match [-all|-bnf|-reverse|-ssnf] "
                             "[-batch IN OUT] [-regex R] [-device D] [-chunk N]\n");
        return 2;
    }
    if (argc > 2 && argv[2][0] >= '0' && argv[2][0] <= '9') {  // main.cpp:11-13: -match N
        int device = 0;
        size_t max_len = size_t(1) << 24;
        const char *bits_path = nullptr;
        for (int i = 3; i + 1 < argc; i++) {
            if (std::strcmp(argv[i], "-device") == 0) device = std::atoi(argv[i + 1]);
            if (std::strcmp(argv[i], "-maxlen") == 0) max_len = std::strtoull(argv[i + 1], nullptr, 10);
            if (std::strcmp(argv[i], "-bits") == 0) bits_path = argv[i + 1];
        }
        return run_configuration_examples_gpu(argv[2], device, max_len, bits_path);
    }
    bool bnf = false, reverse = false, ssnf = false;
    const char *batch_in = nullptr, *batch_out = nullptr;
    std::string regex;
    bool have_regex = false;
    int device = 0;
    uint64_t chunk = uint64_t(256) << 20;
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
    // stdin is read with read(2) so that nothing is buffered away from the token loop below
    std::vector<uint8_t> buf;
    size_t pos = 0;  // first unconsumed byte of buf
    bool eof = false;
    auto fill = [&]() {  // one read(2); false at end of input
        const size_t old = buf.size(), want = size_t(8) << 20;
        buf.resize(old + want);
        const long got = rxm_posix_read(0, buf.data() + old, want);
        buf.resize(old + (got > 0 ? size_t(got) : 0));
        if (got <= 0) eof = true;
        return got > 0;
    };
    if (!have_regex) {  // main.cpp:42-43: cin >> regex
        for (;;) {
            while (pos < buf.size() && is_ws(buf[pos])) pos++;
            size_t e = pos;
            while (e < buf.size() && !is_ws(buf[e])) e++;
            if (e > pos && (e < buf.size() || eof)) {
                regex.assign(reinterpret_cast<const char *>(buf.data()) + pos, e - pos);
                pos = e;
                break;
            }
            if (eof) return 2;
            fill();
        }
    }

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
        bool sane = off[0] == 0 && off[n] == total;  // the header is not trusted: rxm_match_batch copies off[n] bytes
        for (uint64_t i = 0; i < n && sane; i++) sane = off[i] <= off[i + 1];
        if (!sane) {
            std::fprintf(stderr, "diploma_rxm: %s: offsets do not describe %llu bytes\n", batch_in, (unsigned long long)total);
            return 2;
        }
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

    // match.cpp:22-31.  The raw bytes go to the device as they are; pieces are cut at
    // whitespace so that no token is split.  A piece is flushed when `chunk_bytes` have
    // gathered, at end of input, or -- on a terminal -- when a line is complete.
    const bool tty = rxm_posix_isatty(0) != 0;
    const size_t chunk_bytes = size_t(chunk) < (size_t(1) << 16) ? (size_t(1) << 16)
                             : (size_t(chunk) > (size_t(3) << 30) ? (size_t(3) << 30) : size_t(chunk));
    bool done = false;
    while (!done) {
        const bool line = tty && buf.size() > pos && buf.back() == '\n';
        if (!eof && !line && buf.size() - pos < chunk_bytes) {
            fill();
            continue;
        }
        size_t end = buf.size();
        if (!eof) {  // keep the (possibly incomplete) last token for the next piece
            while (end > pos && !is_ws(buf[end - 1])) end--;
            if (end == pos) {  // one token longer than everything read so far
                fill();
                continue;
            }
        }
        if (end - pos > (size_t(3) << 30)) {  // rxm_match_text takes < 4 GiB: cut at whitespace
            end = pos + (size_t(3) << 30);
            while (end > pos && !is_ws(buf[end - 1])) end--;
            if (end == pos) {
                std::fprintf(stderr, "diploma_rxm: a token of more than 3 GiB\n");
                return 3;
            }
        }
        if (end > pos) st = match_piece(h, buf.data() + pos, end - pos, &done);
        if (st != RXM_OK) break;
        pos = end;
        if (pos == buf.size()) {
            buf.clear();
            pos = 0;
        } else if (pos > (size_t(64) << 20)) {
            buf.erase(buf.begin(), buf.begin() + pos);
            pos = 0;
        }
        if (eof && pos == buf.size()) break;
    }
    if (st != RXM_OK) {
        std::fprintf(stderr, "diploma_rxm: match: %s (%s)\n", rxm_strerror(st), rxm_last_cuda_error());
        return 3;
    }
    rxm_free(h);
    return 0;
}
