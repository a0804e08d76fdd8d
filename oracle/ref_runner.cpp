// TEST INFRASTRUCTURE (oracle) -- not part of the product path.
//
// The reference's BENCHMARK ROUTE with its match bits made visible.
// matchers/example_runner.cpp:84-151 (`./diploma -match N`) builds three automata of one expression
// (plain / -bnf / -reverse, :109-111), feeds them CUMULATIVE attack strings (:123) and records only
// how long each match took -- the result of MFA::match is thrown away (:46-51).  This driver runs
// the same loop over the same strings and prints the bits instead of the times:
//     RUNNER <length> <bit plain> <bit bnf> <bit reverse>
// (and, as an extra that the reference's loop does not have, a RUNNER_NS line for the same string
// before its failing suffix is appended -- the strings of the reference's route all end in it).
// The strings come from the reference's OWN generator: pumped_string (example_runner.cpp:15-29) and
// split (:31-44) are linked from matchers/example_runner.cpp compiled where it lies (oracle/Makefile;
// its 5-argument compile() calls are reduced to the 4 arguments regex/regex.h:226 declares by a
// command-line macro -- the same 5th-argument defect main.cpp and match.cpp have at HEAD).
// Everything that decides a bit is the reference's code: Regexp::parse_regexp, Regexp::compile,
// MFA::match.  No timing, no watchdog thread: every string up to -maxlen is matched by all three.
//
//   diploma_ref_runner N [-maxlen L]      reads test/example_N/{regexp,pump}.txt under the cwd
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "regex/regex.h"  // from -I/root/reference
#include "automata.h"

using namespace std;

// defined in the reference's matchers/example_runner.cpp
std::string pumped_string(int n, vector<string> pump_v);
vector<string> split(string str, char separator);

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    const string number = argv[1];
    size_t max_len = 20000;
    for (int i = 2; i + 1 < argc; i++)
        if (strcmp(argv[i], "-maxlen") == 0) max_len = strtoull(argv[i + 1], nullptr, 10);
    fstream regex_file("test/example_" + number + "/regexp.txt");   // example_runner.cpp:85-86
    fstream pump_file("test/example_" + number + "/pump.txt");
    if (!regex_file.is_open() || !pump_file.is_open()) return 2;
    string pump_s, suffix, prefix, regexp_str;
    getline(pump_file, pump_s);                                     // :97-100
    auto pump = split(pump_s, ',');
    getline(pump_file, suffix);
    getline(pump_file, prefix);
    int pump_size = 500;                                            // :102
    getline(regex_file, regexp_str);                                // :104
    Regexp *regexp = Regexp::parse_regexp(regexp_str);              // :106-107
    regexp->is_backref_correct();
    bool is_mfa = true;
    MFA *mfa[3];
    mfa[0] = static_cast<MFA *>(regexp->compile(is_mfa, false, false, true));   // :109-111
    mfa[1] = static_cast<MFA *>(regexp->compile(is_mfa, false, true, true));
    mfa[2] = static_cast<MFA *>(regexp->compile(is_mfa, true, true, true));
    cout.flush();
    int count = 0;
    size_t len = prefix.length() + pump_size + suffix.length();     // :118
    while (len < max_len) {                                         // :120 (no timeouts here)
        // (extra, not in the reference's loop: the same string before its failing suffix is appended)
        string no_suffix = prefix + pumped_string(pump_size, pump);
        string input_str = prefix.append(pumped_string(pump_size, pump)).append(suffix);   // :123
        len = input_str.length();
        if (len > max_len) break;
        pump_size += pump_size;                                     // :125
        printf("RUNNER %zu", len);
        for (int a = 0; a < 3; a++) printf(" %d", mfa[a]->match(input_str) ? 1 : 0);
        printf("\n");
        printf("RUNNER_NS %zu", no_suffix.length());
        for (int a = 0; a < 3; a++) printf(" %d", mfa[a]->match(no_suffix) ? 1 : 0);
        printf("\n");
        fflush(stdout);
        count++;                                                    // :140
        if (count % 10 == 0) pump_size *= 2;                        // :143-144
    }
    return 0;
}
