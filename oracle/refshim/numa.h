/* stub of <numa.h> for the non-NUMA build (SPX_USE_NUMA = 0): declarations only, never called */
#ifndef REFSHIM_NUMA_H
#define REFSHIM_NUMA_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
struct bitmask;
void *numa_alloc_onnode(size_t size, int node);
void *numa_alloc_interleaved(size_t size);
void *numa_alloc_local(size_t size);
void *numa_realloc(void *old_addr, size_t old_size, size_t new_size);
void numa_free(void *start, size_t size);
int numa_node_of_cpu(int cpu);
int numa_num_configured_nodes(void);
#ifdef __cplusplus
}
#endif
#endif
