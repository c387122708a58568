/* TEST INFRASTRUCTURE.  Stand-in for the one internal header the reference's test/src/CsxCheck.hpp pulls in when it is
 * compiled as C (by test/src/sparsex_test.c): the names that header needs, on top of this repository's public headers.
 * Lets tests/test_reference_examples.py compile the reference's own API test program unchanged. */
#ifndef SPXB_TEST_SPMMT_STANDIN
#define SPXB_TEST_SPMMT_STANDIN
#include <sparsex/sparsex.h>
typedef struct vector_struct vector_t;   /* include/sparsex/internals/Vector.hpp:30-35 of the reference */
#ifndef SPX_BEGIN_C_DECLS__
#ifdef __cplusplus
#define SPX_BEGIN_C_DECLS__ extern "C" {
#define SPX_END_C_DECLS__ }
#else
#define SPX_BEGIN_C_DECLS__
#define SPX_END_C_DECLS__
#endif
#endif
#endif
