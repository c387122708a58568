/* sparsex/sparsex.h — umbrella header of the drop-in SparseX API implemented
 * by the B200 CSX SpMV engine (libsparsex_b200.so). */
#ifndef SPARSEX_SPARSEX_H
#define SPARSEX_SPARSEX_H

#include <sparsex/common.h>
#include <sparsex/error.h>
#include <sparsex/matvec.h>
#include <sparsex/timing.h>
#include <sparsex/types.h>

#endif /* SPARSEX_SPARSEX_H */
