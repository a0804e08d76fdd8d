// Link-time shim for the reference front end (parse / bnf / reverse / automaton
// builders), which the product keeps on the CPU unmodified.
//
// Automata::draw (automata.cpp:68-82) and MFA::draw (mfa.cpp:63-77) are declared
// `bool` but fall off the end; g++ >= 13 turns that into a trap and
// Regexp::compile calls them unconditionally (regex/regex.cpp:287,294,311,331).
// The front-end binaries are linked with -Wl,--wrap=<mangled name> so those
// calls land here: same side effects (makeDOTFile), plus the missing return.
#include <string>

#include "automata.h"  // reference header, via -I<reference root>

extern "C" bool
__wrap__ZN8Automata4drawERKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEE(
    Automata *self, const std::string &filename) {
    self->makeDOTFile(filename);
    return true;
}

extern "C" bool
__wrap__ZN3MFA4drawERKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEE(
    MFA *self, const std::string &filename) {
    self->makeDOTFile(filename);
    return true;
}

#ifndef RXM_FRONT_STOCK_MALLOC
// Canonical state numbering.  The flattening stage numbers states by the address
// rank of the reference's Node objects, because that is the order in which the
// reference's std::set<Node*> / std::set<MemoryState> visit them.  With glibc
// malloc that rank depends on which freed chunks happen to be recycled while the
// expression is parsed and normalised, i.e. on process history.  The front-end
// binaries therefore replace global operator new with a never-reuse bump
// allocator for SMALL objects: address order == allocation order, so the tables
// are a pure function of (expression, flags).  (regex/bnf.cpp:222 releases a
// new'd object with ::free(); -Wl,--wrap=free routes that here.)
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <sys/mman.h>

namespace {
constexpr size_t kArenaBytes = size_t(1) << 36;  // 64 GiB of address space, committed lazily
constexpr size_t kSmallMax = 256;
char *g_base = nullptr, *g_cur = nullptr;

void *bump(size_t n) {
    if (!g_base) {
        void *p = mmap(nullptr, kArenaBytes, PROT_READ | PROT_WRITE,
                       MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) abort();
        g_base = g_cur = static_cast<char *>(p);
    }
    n = n ? (n + 15) & ~size_t(15) : 16;
    if (size_t(g_cur - g_base) + n > kArenaBytes) {
        fputs("rxm front end: bump arena exhausted\n", stderr);
        abort();
    }
    char *r = g_cur;
    g_cur += n;
    return r;
}
bool in_arena(const void *p) { return g_base && p >= g_base && p < g_base + kArenaBytes; }
void *front_new(size_t n) {
    if (n <= kSmallMax) return bump(n);
    // zero-filled like the arena: the reference reads members it never initialises (`bool is_read`,
    // `Regexp* reference_to`, regex/regex.h:85; first read at regex/bnf.cpp:49,63), and a Regexp is
    // larger than kSmallMax -- with recycled malloc chunks bnf() then builds a different expression
    // from one process to the next.  Zero is what a fresh heap gives the reference.
    void *p = calloc(1, n);
    if (!p) abort();
    return p;
}
}  // namespace

extern "C" void __real_free(void *);
extern "C" void __wrap_free(void *p) {
    if (!in_arena(p)) __real_free(p);
}
static void front_delete(void *p) {
    if (p && !in_arena(p)) __real_free(p);
}
void *operator new(size_t n) { return front_new(n); }
void *operator new[](size_t n) { return front_new(n); }
void *operator new(size_t n, const std::nothrow_t &) noexcept { return front_new(n); }
void *operator new[](size_t n, const std::nothrow_t &) noexcept { return front_new(n); }
void operator delete(void *p) noexcept { front_delete(p); }
void operator delete[](void *p) noexcept { front_delete(p); }
void operator delete(void *p, size_t) noexcept { front_delete(p); }
void operator delete[](void *p, size_t) noexcept { front_delete(p); }
#endif
