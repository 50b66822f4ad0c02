// shim.cpp — the few symbols the reference objects expect from parts of the reference that are NOT built here
// (TEST INFRASTRUCTURE; see oracle/ref/README.md).  Nothing here is on a numerical path.
#include <cstddef>
#include <cstdlib>
#include <new>

// EASTL expects the application to provide these two allocation operators (EASTL/allocator.h; the reference provides them in
// its EASTL fork's allocator sources when built through CMake with EASTL_USER_DEFINED_ALLOCATOR).
void *operator new[](size_t size, const char *, int, unsigned, const char *, int) { return ::operator new[](size); }
void *operator new[](size_t size, size_t alignment, size_t, const char *, int, unsigned, const char *, int) {
    return ::operator new[](size, std::align_val_t{alignment});
}

