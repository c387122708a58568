/* sparsex/error.h — error codes and the error-handler hook of the drop-in API.
 * Codes and handler signature follow include/sparsex/error.h:34-147 of SparseX. */
#ifndef SPARSEX_ERROR_H
#define SPARSEX_ERROR_H

#include <stdarg.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPX_FAILURE             -1
#define SPX_SUCCESS             0

#define SPX_ERR_MIN_VALUE       1
#define SPX_ERR_ARG_INVALID     2   /* invalid argument */
#define SPX_ERR_FILE            3   /* generic file error */
#define SPX_ERR_INPUT_MAT       4   /* input matrix wasn't properly created */
#define SPX_ERR_TUNED_MAT       5   /* tuned matrix wasn't properly created */
#define SPX_ERR_VEC             6   /* vector creation failed */
#define SPX_ERR_PART            7   /* partitioning object wasn't properly created */
#define SPX_ERR_PERM            8   /* error in permutation */
#define SPX_ERR_DIM             9   /* incompatible matrix and vector dimensions */
#define SPX_ERR_VEC_DIM         10  /* incompatible vector dimension */
#define SPX_ERR_ENTRY_NOT_FOUND 11  /* matrix entry not found */
#define SPX_OUT_OF_BOUNDS       12  /* index out of bounds */
#define SPX_ERR_SYSTEM          15
#define SPX_ERR_FILE_OPEN       16
#define SPX_ERR_FILE_READ       17
#define SPX_ERR_FILE_WRITE      18
#define SPX_ERR_MEM_ALLOC       19
#define SPX_ERR_MEM_FREE        20
#define SPX_ERR_MAX_VALUE       21

#define SPX_WARN_CSXFILE        22
#define SPX_WARN_TUNING_OPT     23
#define SPX_WARN_RUNTIME_OPT    24
#define SPX_WARN_REORDER        25
#define SPX_WARN_ENTRY_NOT_SET  26
#define SPX_WARN_MAX_VALUE      27

typedef int spx_error_t;
typedef void (*spx_errhandler_t)(spx_error_t, const char *, unsigned long, const char *, const char *, ...);

#define SETERROR_0(code) spx_err_get_handler()(code, __FILE__, __LINE__, __func__, NULL)
#define SETERROR_1(code, message) spx_err_get_handler()(code, __FILE__, __LINE__, __func__, message)
#define SETWARNING(code) spx_err_get_handler()(code, __FILE__, __LINE__, __func__, NULL)

/* Default handler: prints "[ERROR|WARNING] in function() ...: message" to stderr;
 * system errors (codes 16-20) terminate the process like the reference does. */
void err_handle(spx_error_t code, const char *sourcefile, unsigned long lineno, const char *function,
                const char *fmt, ...);
spx_errhandler_t spx_err_get_handler(void);
void spx_err_set_handler(spx_errhandler_t new_handler);

#ifdef __cplusplus
}
#endif
#endif /* SPARSEX_ERROR_H */
