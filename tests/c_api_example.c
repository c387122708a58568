/* Drop-in check: a C program written against <sparsex/sparsex.h> the way the reference's
 * src/examples/csr_example.c and test/src/sparsex_test.c use it.  Built by tests/test_c_program.py. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <sparsex/sparsex.h>

int main(int argc, char **argv)
{
    /* 2-D 5-point Laplacian on a g x g grid as zero-based CSR */
    int g = argc > 1 ? atoi(argv[1]) : 64, n = g * g, nnz = 0;
    spx_index_t *rowptr = malloc((n + 1) * sizeof(*rowptr));
    spx_index_t *colind = malloc(5 * (size_t) n * sizeof(*colind));
    spx_value_t *values = malloc(5 * (size_t) n * sizeof(*values));
    for (int i = 0; i < n; i++) {
        int gx = i % g, gy = i / g;
        rowptr[i] = nnz;
        if (gy > 0) { colind[nnz] = i - g; values[nnz++] = -1.0; }
        if (gx > 0) { colind[nnz] = i - 1; values[nnz++] = -1.0; }
        colind[nnz] = i; values[nnz++] = 4.0;
        if (gx < g - 1) { colind[nnz] = i + 1; values[nnz++] = -1.0; }
        if (gy < g - 1) { colind[nnz] = i + g; values[nnz++] = -1.0; }
    }
    rowptr[n] = nnz;

    spx_init();
    spx_log_error_console();
    if (argc > 2) spx_option_set("spx.matrix.symmetric", argv[2]);
    spx_option_set("spx.preproc.xform", "all");
    spx_input_t *input = spx_input_load_csr(rowptr, colind, values, n, n);
    if (input == SPX_INVALID_INPUT) return 2;
    spx_matrix_t *A = spx_mat_tune(input);
    if (A == SPX_INVALID_MAT) return 3;
    spx_input_destroy(input);
    if (spx_mat_get_nrows(A) != n || spx_mat_get_nnz(A) != nnz) return 4;

    spx_partition_t *parts = spx_mat_get_partition(A);
    spx_vector_t *x = spx_vec_create_random(n, parts);
    spx_vector_t *y = spx_vec_create(n, parts);
    spx_value_t *ybuf = malloc(n * sizeof(*ybuf)), *ytuned = NULL;
    spx_vector_t *y2 = spx_vec_create_from_buff(ybuf, &ytuned, n, parts, SPX_VEC_AS_IS);
    const spx_value_t alpha = 0.8, beta = 0.42;

    if (spx_matvec_mult(alpha, A, x, y) != SPX_SUCCESS) return 5;        /* managed vectors */
    if (spx_matvec_mult(alpha, A, x, y2) != SPX_SUCCESS) return 6;       /* user buffer for y */
    double maxerr = 0, maxmix = 0;
    for (int i = 0; i < n; i++) {
        double r = 0;
        for (int k = rowptr[i]; k < rowptr[i + 1]; k++) r += values[k] * x->elements[colind[k]];
        maxerr = fmax(maxerr, fabs(y->elements[i] - alpha * r));
        maxmix = fmax(maxmix, fabs(ytuned[i] - y->elements[i]));
    }
    /* y <- alpha*A*x + beta*y */
    if (spx_matvec_kernel(alpha, A, x, beta, y) != SPX_SUCCESS) return 7;
    double maxerr2 = 0;
    for (int i = 0; i < n; i++) {
        double r = 0;
        for (int k = rowptr[i]; k < rowptr[i + 1]; k++) r += values[k] * x->elements[colind[k]];
        maxerr2 = fmax(maxerr2, fabs(y->elements[i] - (alpha * r + beta * ytuned[i])));
    }
    /* dimension mismatch must fail through the handler, not crash */
    spx_vector_t *bad = spx_vec_create(n + 1, parts);
    int rc_bad = spx_matvec_mult(alpha, A, bad, y);
    printf("n=%d nnz=%d maxerr=%.3e maxerr_kernel=%.3e buff_vs_managed=%.3e bad_dim_rc=%d\n", n, nnz, maxerr, maxerr2,
           maxmix, rc_bad);
    spx_vec_destroy(bad); spx_vec_destroy(x); spx_vec_destroy(y); spx_vec_destroy(y2);
    spx_partition_destroy(parts);
    spx_mat_destroy(A);
    spx_finalize();
    free(rowptr); free(colind); free(values); free(ybuf);
    /* CSX-Sym adds with fp64 red operations: summation order, hence the last bits, may differ between calls */
    return (maxerr < 1e-12 && maxerr2 < 1e-12 && maxmix < 1e-12 && rc_bad == SPX_FAILURE) ? 0 : 1;
}
