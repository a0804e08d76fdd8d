// TEST INFRASTRUCTURE (oracle) -- not part of the product path.
//
// Link-time shims that let the UNMODIFIED reference sources under
// /root/reference run under g++ >= 13 without editing them:
//
//  * Automata::draw (automata.cpp:68-82) and MFA::draw (mfa.cpp:63-77) are
//    declared `bool` but fall off the end; g++ 13 turns that into a trap, and
//    Regexp::compile calls draw() unconditionally (regex/regex.cpp:287,294,311,331).
//    The link uses `-Wl,--wrap=<mangled draw>` so regex.cpp's calls land here;
//    we do what the body does (makeDOTFile) and return true.
//
//  * (bump variant only, -DRXM_ORACLE_BUMP) global operator new/delete are
//    replaced by a never-reuse bump allocator so that "pointer order ==
//    allocation order".  The reference orders std::set<MemoryState> by
//    Variable* ADDRESS (automata.h:12-13), so with glibc malloc its answers
//    depend (rarely) on heap reuse; the bump allocator makes the reference a
//    pure function of (regex, flags, string).  regex/bnf.cpp:222 calls
//    ::free() on new'd memory, so `-Wl,--wrap=free` routes free() here and
//    arena pointers are ignored.
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <string>
#include <sys/mman.h>

#include "automata.h"  // from -I/root/reference

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

#ifdef RXM_ORACLE_BUMP
namespace {
// 1 TiB of lazily committed address space: enough for any oracle run (the
// reference leaks ~O(len * edges) cells per string, mfa.cpp:107-114).
constexpr size_t ARENA_BYTES = size_t(1) << 40;
char *arena_base = nullptr;
char *arena_cur = nullptr;

inline void arena_init() {
    void *p = mmap(nullptr, ARENA_BYTES, PROT_READ | PROT_WRITE,
                   MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) {
        fputs("oracle bump arena: mmap failed\n", stderr);
        abort();
    }
    arena_base = arena_cur = static_cast<char *>(p);
}

inline void *arena_alloc(size_t n) {
    if (!arena_base) arena_init();
    n = (n + 15) & ~size_t(15);
    if (n == 0) n = 16;
    char *r = arena_cur;
    arena_cur += n;
    if (size_t(arena_cur - arena_base) > ARENA_BYTES) {
        fputs("oracle bump arena exhausted\n", stderr);
        abort();
    }
    return r;
}
inline bool in_arena(const void *p) {
    return arena_base && p >= arena_base && p < arena_base + ARENA_BYTES;
}
}  // namespace

extern "C" void __real_free(void *);

// Only SMALL objects need "address order == allocation order": Variable (48 B,
// variable.h:8-41), Node/MemoryNode, Edge, std::set / std::map tree nodes.
// Large blocks are the by-value copies of the input string
// (mfa.cpp:136,203,215); they never take part in an ordering, so they go to
// malloc and ARE freed -- otherwise one 64 K-char string would pin ~40 GB.
constexpr size_t SMALL_MAX = 256;
static inline void *oracle_new(size_t n) {
    if (n <= SMALL_MAX) return arena_alloc(n);
    // zero-filled like the arena: the reference reads members it never initialises (`bool is_read`,
    // `Regexp* reference_to`, regex/regex.h:85; first read at regex/bnf.cpp:49,63), and a Regexp is
    // larger than SMALL_MAX -- with recycled malloc chunks bnf() then builds a different expression
    // from one process to the next.  Zero is what a fresh heap gives the reference.
    void *p = calloc(1, n);
    if (!p) abort();
    return p;
}
static inline void oracle_delete(void *p) {
    if (p && !in_arena(p)) __real_free(p);
}

void *operator new(size_t n) { return oracle_new(n); }
void *operator new[](size_t n) { return oracle_new(n); }
void *operator new(size_t n, const std::nothrow_t &) noexcept { return oracle_new(n); }
void *operator new[](size_t n, const std::nothrow_t &) noexcept { return oracle_new(n); }
void operator delete(void *p) noexcept { oracle_delete(p); }
void operator delete[](void *p) noexcept { oracle_delete(p); }
void operator delete(void *p, size_t) noexcept { oracle_delete(p); }
void operator delete[](void *p, size_t) noexcept { oracle_delete(p); }

extern "C" void __wrap_free(void *p) {
    if (in_arena(p)) return;
    __real_free(p);
}

// Never rewound.  The driver may ask how much was consumed (for sizing subsamples).
extern "C" size_t rxm_oracle_arena_used() { return size_t(arena_cur - arena_base); }
#endif
