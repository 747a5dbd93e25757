// TEST INFRASTRUCTURE — part of oracle/_ref/libref.so only.
//
// ORBextractor::DistributeOctTree sorts pair<int size, ExtractorNode* node> (ORBextractor.cc:684) and walks it
// from the back, so among quadtree nodes that hold the same number of keypoints the one at the HIGHEST heap
// address is split first.  Under glibc malloc that order depends on the allocator's free lists.  This file
// replaces operator new inside libref.so (linked -Bsymbolic-functions, so nothing outside the library is
// affected) with a per-thread bump arena that is active only while the harness runs the reference extractor:
// addresses ascend with creation order, nothing is reused, and the sort becomes reproducible —
// "latest created first" among ties, which is what oracle/orb_oracle.cpp and the CUDA k_octree define.
#include <sys/mman.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <new>

namespace refarena {
constexpr size_t kArenaBytes = size_t(1) << 30;   // virtual; pages are touched on demand and reused per call
struct Arena { char* base = nullptr; size_t used = 0; bool active = false; };
static thread_local Arena t_arena;
static char* g_lo = nullptr;   // all arenas are carved from one reservation, so delete can range-check
static char* g_hi = nullptr;
static size_t g_next = 0;
constexpr size_t kMaxThreads = 256;

static void reserve_once() {
  static bool done = [] {
    void* p = mmap(nullptr, kArenaBytes * kMaxThreads, PROT_READ | PROT_WRITE,
                   MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) { std::perror("refarena mmap"); std::abort(); }
    g_lo = (char*)p; g_hi = g_lo + kArenaBytes * kMaxThreads;
    return true;
  }();
  (void)done;
}

void begin() {
  // REF_ARENA_OFF=1: leave the reference on glibc malloc (used by tests to show the tie-break really is address order)
  static const bool off = std::getenv("REF_ARENA_OFF") != nullptr;
  if (off) return;
  Arena& a = t_arena;
  if (!a.base) {
    reserve_once();
    size_t slot = __atomic_fetch_add(&g_next, 1, __ATOMIC_RELAXED);
    if (slot >= kMaxThreads) { std::fprintf(stderr, "refarena: too many threads\n"); std::abort(); }
    a.base = g_lo + slot * kArenaBytes;
  }
  a.used = 0;
  a.active = true;
}
void end() { t_arena.active = false; }

static inline void* alloc(size_t n) {
  Arena& a = t_arena;
  if (a.active) {
    size_t off = (a.used + 15) & ~size_t(15);
    if (off + n > kArenaBytes) { std::fprintf(stderr, "refarena: arena exhausted\n"); std::abort(); }
    a.used = off + n;
    return a.base + off;
  }
  void* p = std::malloc(n ? n : 1);
  if (!p) throw std::bad_alloc();
  return p;
}
static inline void dealloc(void* p) {
  if (!p) return;
  if ((char*)p >= g_lo && (char*)p < g_hi) return;   // arena memory is recycled wholesale by begin()
  std::free(p);
}
}  // namespace refarena

void* operator new(size_t n) { return refarena::alloc(n); }
void* operator new[](size_t n) { return refarena::alloc(n); }
void operator delete(void* p) noexcept { refarena::dealloc(p); }
void operator delete[](void* p) noexcept { refarena::dealloc(p); }
void operator delete(void* p, size_t) noexcept { refarena::dealloc(p); }
void operator delete[](void* p, size_t) noexcept { refarena::dealloc(p); }
