/*
 * TEST INFRASTRUCTURE ONLY (oracle build).  Minimal stand-in for <numa.h> so the
 * unmodified reference sources under /root/reference compile on a box without
 * libnuma-dev.  numa_available() reports "no NUMA", which makes the reference take
 * its own malloc/realloc/free branch (reference src/pcsr/PCSR.cpp:776,787-790) and
 * build a single NUMA domain (reference src/thread_pool_pppcsr/thread_pool_pppcsr.cpp:24).
 * Call sites that need these symbols: reference src/pcsr/PCSR.cpp:76-79,257-274,
 * 306-317,781-786,844-846 and src/thread_pool_pppcsr/thread_pool_pppcsr.cpp:24,57-58.
 */
#ifndef PPCSR_ORACLE_NUMA_STUB_H
#define PPCSR_ORACLE_NUMA_STUB_H
#include <stdlib.h>
#ifdef __cplusplus
extern "C" {
#endif
static inline int numa_available(void) { return -1; }
static inline int numa_max_node(void) { return 0; }
static inline int numa_run_on_node(int node) { (void)node; return 0; }
static inline void *numa_alloc_onnode(size_t size, int node) { (void)node; return malloc(size); }
static inline void *numa_realloc(void *p, size_t old_size, size_t new_size) { (void)old_size; return realloc(p, new_size); }
static inline void numa_free(void *p, size_t size) { (void)size; free(p); }
#ifdef __cplusplus
}
#endif
#endif
