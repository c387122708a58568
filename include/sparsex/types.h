/* sparsex/types.h — index and value types of the drop-in API.
 * Same ABI as the reference build defaults (include/sparsex/types.h:25-35,
 * configure.ac:98,111): spx_index_t = int, spx_value_t = double. */
#ifndef SPARSEX_TYPES_H
#define SPARSEX_TYPES_H

typedef int spx_index_t;
typedef double spx_value_t;

#endif /* SPARSEX_TYPES_H */
