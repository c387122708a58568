/* TEST INFRASTRUCTURE.  check_result() for the reference's API test program (test/src/sparsex_test.c, compiled unchanged
 * from /root/reference): what test/src/CsxCheck.hpp:85-107 + CsxCheck.cpp:23-54 do there with the reference's internals —
 * read the MatrixMarket file again, multiply it as plain CSR on one thread, scale by alpha, compare with the library's
 * result through spx_vec_compare (relative 1e-6, Vector.cpp:396-413) and exit(1) on a mismatch — written on this
 * repository's public API only.  The file format is the reference's (Mmf.hpp:331-478): optional "%%MatrixMarket matrix
 * coordinate real general|symmetric [0-base|1-base] [row|column]" banner, size line, "row col value" lines. */
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <sparsex/sparsex.h>

typedef struct vector_struct vector_t;

void check_result(vector_t *result, double alpha, vector_t *x, char *matrix_file) {
  FILE *f = fopen(matrix_file, "r");
  if (!f) { fprintf(stderr, "check_result: cannot open %s\n", matrix_file); exit(1); }
  char line[1024];
  int symmetric = 0, zero_based = 0, have = 0;
  long nr = 0, nc = 0, nnz = 0;
  while (fgets(line, sizeof line, f)) {
    if (line[0] == '%') {
      for (char *p = line; *p; p++) *p = (char)tolower((unsigned char)*p);
      if (strstr(line, "symmetric")) symmetric = 1;
      if (strstr(line, "0-base")) zero_based = 1;
      continue;
    }
    if (sscanf(line, "%ld %ld %ld", &nr, &nc, &nnz) == 3) { have = 1; break; }
  }
  if (!have) { fprintf(stderr, "check_result: no size line in %s\n", matrix_file); exit(1); }
  double *y = (double *)calloc((size_t)(nr ? nr : 1), sizeof(double));
  printf("Checking... ");
  fflush(stdout);
  for (long k = 0; k < nnz; k++) {
    long r, c;
    double v;
    if (fscanf(f, "%ld %ld %lf", &r, &c, &v) != 3) { fprintf(stderr, "check_result: short file\n"); exit(1); }
    if (!zero_based) { r--; c--; }
    y[r] += v * x->elements[c];
    if (symmetric && r != c) y[c] += v * x->elements[r];
  }
  fclose(f);
  spx_vector_t *y_csr = spx_vec_create_from_buff(y, NULL, (size_t)nr, NULL, SPX_VEC_AS_IS);
  for (long i = 0; i < nr; i++) y[i] *= alpha;
  if (spx_vec_compare(y_csr, result) < 0) exit(1);
  printf("Check Passed\n");
  spx_vec_destroy(y_csr);
  free(y);
}
