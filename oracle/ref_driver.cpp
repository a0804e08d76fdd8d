// TEST INFRASTRUCTURE (oracle) -- not part of the product path.
//
// Driver around the UNMODIFIED reference library objects (compiled from
// /root/reference by oracle/Makefile).  It stands in for main.cpp:9-45 +
// matchers/match.cpp:10-32, which do not compile at the reference's HEAD
// (they pass a 5th `use_log` argument that regex/regex.h:226,233 does not
// declare).  Everything that decides a match bit is the reference's own code:
//   Regexp::parse_regexp   regex/parser.cpp:8
//   Regexp::compile        regex/regex.cpp:266-343   (engine selection)
//   MFA::match             mfa.cpp:215-236
//   Automata::match        automata.cpp:177-210
//
// Modes
//   diploma_ref [-match] [-all|-bnf|-reverse|-ssnf ...]
//       stdin protocol of match.cpp: first token = regex, then one result line
//       ("0"/"1") per whitespace-delimited token until the token `exit`
//       (or EOF -- the reference would spin forever on EOF, match.cpp:23-31).
//   diploma_ref [flags] -regex R -batch IN.rxmb OUT.bits [-range LO HI]
//       match strings LO..HI-1 of a binary batch (format below) and write one
//       byte (0/1) per string; prints "ORACLE_TIME <seconds> <nstrings> <nchars>"
//       on stderr, timing only the match loop (no compile, no I/O).
//
// Batch file: "RXMBATCH" | u64 n | u64 total | u64 offsets[n+1] | u8 chars[total]
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "regex/regex.h"  // from -I/root/reference
#include "automata.h"

using namespace std;

static bool read_batch(const char *path, vector<uint64_t> &off, vector<char> &chars) {
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    char magic[8];
    uint64_t n = 0, total = 0;
    bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, "RXMBATCH", 8) == 0 &&
              fread(&n, 8, 1, f) == 1 && fread(&total, 8, 1, f) == 1;
    if (ok) {
        off.resize(n + 1);
        chars.resize(total);
        ok = fread(off.data(), 8, n + 1, f) == n + 1 &&
             (total == 0 || fread(chars.data(), 1, total, f) == total);
    }
    fclose(f);
    return ok;
}

int main(int argc, char **argv) {
    bool bnf = false, reverse = false, ssnf = false;
    const char *batch_in = nullptr, *batch_out = nullptr;
    string regex_arg;
    bool have_regex = false;
    long lo = 0, hi = -1;
    int first_flag = 1;
    if (argc > 1 && strcmp(argv[1], "-match") == 0) first_flag = 2;
    // main.cpp:28-40: `-all` counts only as the first flag; -reverse implies -bnf.
    if (argc > first_flag && strcmp(argv[first_flag], "-all") == 0) bnf = reverse = ssnf = true;
    for (int i = first_flag; i < argc; i++) {
        string a = argv[i];
        if (a == "-bnf") bnf = true;
        else if (a == "-reverse") { reverse = true; bnf = true; }
        else if (a == "-ssnf") ssnf = true;
        else if (a == "-regex" && i + 1 < argc) { regex_arg = argv[++i]; have_regex = true; }
        else if (a == "-batch" && i + 2 < argc) { batch_in = argv[++i]; batch_out = argv[++i]; }
        else if (a == "-range" && i + 2 < argc) { lo = atol(argv[++i]); hi = atol(argv[++i]); }
    }

    string regex;
    if (have_regex) regex = regex_arg;
    else cin >> regex;  // main.cpp:42-43

    Regexp *regexp = Regexp::parse_regexp(regex);                  // match.cpp:12
    bool is_mfa = false;
    Automata *automata = regexp->compile(is_mfa, reverse, bnf, ssnf);  // match.cpp:15
    MFA *mfa = is_mfa ? static_cast<MFA *>(automata) : nullptr;    // match.cpp:17-19
    cout.flush();

    if (batch_in) {
        vector<uint64_t> off;
        vector<char> chars;
        if (!read_batch(batch_in, off, chars)) {
            fprintf(stderr, "cannot read batch %s\n", batch_in);
            return 2;
        }
        long n = long(off.size()) - 1;
        if (hi < 0 || hi > n) hi = n;
        if (lo < 0) lo = 0;
        vector<unsigned char> bits(hi > lo ? hi - lo : 0);
        uint64_t nchars = 0;
        auto t0 = chrono::steady_clock::now();
        for (long i = lo; i < hi; i++) {
            string text(chars.data() + off[i], chars.data() + off[i + 1]);
            nchars += text.size();
            bool m = mfa ? mfa->match(text) : automata->match(text);  // match.cpp:25-28
            bits[i - lo] = m ? 1 : 0;
        }
        auto t1 = chrono::steady_clock::now();
        FILE *f = fopen(batch_out, "wb");
        if (!f) return 2;
        fwrite(bits.data(), 1, bits.size(), f);
        fclose(f);
        fprintf(stderr, "ORACLE_TIME %.6f %ld %llu\n",
                chrono::duration<double>(t1 - t0).count(), hi - lo, (unsigned long long)nchars);
        return 0;
    }

    string text;
    while (cin >> text) {           // match.cpp:22-31 (+ EOF check)
        if (text == "exit") break;
        bool m = mfa ? mfa->match(text) : automata->match(text);
        cout << m << endl;
    }
    return 0;
}
