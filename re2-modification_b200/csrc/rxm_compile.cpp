// rxm_compile -- front-end tool: expression (+ -match flags) -> table text.
//
// The reference's own code parses, normalises, reverses and builds the
// automaton (Regexp::parse_regexp regex/parser.cpp:8, Regexp::compile
// regex/regex.cpp:266-343); the flattening stage (rxm_flatten.hpp) turns the
// result into rxm_tables, printed in the text form of rxm_tables_format.
//
//   rxm_compile [-all|-bnf|-reverse|-ssnf ...] -regex R [-o FILE]
//
// Flag semantics are main.cpp:28-40 (-all only as the first flag; -reverse
// implies -bnf).  compile() prints its banners to stdout and writes .dot files
// into the cwd (regex.cpp:272-331); here the banners are forwarded to stderr
// and the .dot files go to a scratch directory.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <chrono>
#include <filesystem>
#include <vector>

#include "regex/regex.h"  // reference header

#include "rxm_flatten.hpp"

int main(int argc, char **argv) {
    bool bnf = false, reverse = false, ssnf = false;
    std::string regex, out_path;
    bool have_regex = false;
    int first_flag = 1;
    if (argc > 1 && std::strcmp(argv[1], "-match") == 0) first_flag = 2;
    if (argc > first_flag && std::strcmp(argv[first_flag], "-all") == 0) bnf = reverse = ssnf = true;
    for (int i = first_flag; i < argc; i++) {
        std::string a = argv[i];
        if (a == "-bnf") bnf = true;
        else if (a == "-reverse") { reverse = true; bnf = true; }
        else if (a == "-ssnf") ssnf = true;
        else if (a == "-regex" && i + 1 < argc) { regex = argv[++i]; have_regex = true; }
        else if (a == "-o" && i + 1 < argc) out_path = argv[++i];
    }
    if (!have_regex && !(std::cin >> regex)) {
        std::fprintf(stderr, "rxm_compile: no expression\n");
        return 2;
    }
    // (no <unistd.h> here: the reference's `enum MemoryAction {open, close}`,
    //  edge.h:29-32, collides with the POSIX declarations)
    namespace fs = std::filesystem;
    std::error_code ec;
    const fs::path cwd = fs::current_path();
    const fs::path scratch =
        fs::temp_directory_path() /
        ("rxm_compile_" +
         std::to_string(std::chrono::steady_clock::now().time_since_epoch().count()) + "_" +
         std::to_string(reinterpret_cast<uintptr_t>(&regex) & 0xffffff));
    fs::create_directories(scratch, ec);
    fs::current_path(scratch, ec);
    if (ec) return 2;

    std::ostringstream banners;
    std::streambuf *old = std::cout.rdbuf(banners.rdbuf());
    Regexp *re = Regexp::parse_regexp(regex);
    bool is_mfa = false;
    Automata *a = re->compile(is_mfa, reverse, bnf, ssnf);
    std::cout.rdbuf(old);
    std::cerr << banners.str();

    rxm::HostTables h;
    std::string err;
    int st = rxm::flatten(a, is_mfa, h, &err);
    fs::current_path(cwd, ec);
    fs::remove_all(scratch, ec);
    if (st != RXM_OK) {
        std::fprintf(stderr, "rxm_compile: %s (%s)\n", rxm_strerror(st), err.c_str());
        return 3;
    }
    rxm_tables t = h.view();
    size_t need = 0;
    if (rxm_tables_format(&t, nullptr, 0, &need) != RXM_OK) return 3;
    std::vector<char> buf(need);
    if (rxm_tables_format(&t, buf.data(), buf.size(), &need) != RXM_OK) return 3;
    if (out_path.empty()) {
        std::fputs(buf.data(), stdout);
    } else {
        FILE *f = std::fopen(out_path.c_str(), "w");
        if (!f) return 2;
        std::fputs(buf.data(), f);
        std::fclose(f);
    }
    return 0;
}
