/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * Source generator for the reference's SpMV kernels.  SparseX builds one C
 * translation unit per partition from the text templates in
 * /root/reference/src/templates/ and JIT-compiles it with Clang
 * (CsxJit.hpp:359-732, TemplateText.cpp:40-80).  This file restates that
 * assembly step — which template per unit, the new-row / next-x / body hooks,
 * ${key} substitution with unset keys becoming empty — so that gcc can stand
 * in for the JIT.  The template and header TEXT is not part of this
 * repository: oracle/build_ref.py embeds it, read from where it lies under
 * /root/reference, into oracle/_ref/libcsxref_gen.so (git-ignored).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

/* provided by the generated resource file (build_ref.py) */
struct csxref_resource { const char *name; const char *text; };
extern const struct csxref_resource csxref_templates[];
extern const struct csxref_resource csxref_headers[];

static const char *find_template(const char *name) {
  for (int i = 0; csxref_templates[i].name; i++)
    if (!strcmp(csxref_templates[i].name, name)) return csxref_templates[i].text;
  return NULL;
}

struct buf { char *p; size_t n, cap; };
static void put(struct buf *b, const char *s, size_t len) {
  if (b->n + len + 1 > b->cap) {
    b->cap = (b->n + len + 1) * 2;
    b->p = (char *)realloc(b->p, b->cap);
  }
  memcpy(b->p + b->n, s, len);
  b->n += len;
  b->p[b->n] = 0;
}
static void puts_(struct buf *b, const char *s) { put(b, s, strlen(s)); }

/* TemplateText::DoSubstitute: every ${word} is replaced by its value, or by "" when unset */
static void substitute(struct buf *out, const char *tmpl, const char **keys, const char **vals, int nkeys) {
  const char *s = tmpl;
  while (*s) {
    if (s[0] == '$' && s[1] == '{') {
      const char *e = s + 2;
      while ((*e >= 'a' && *e <= 'z') || (*e >= 'A' && *e <= 'Z') || (*e >= '0' && *e <= '9') || *e == '_') e++;
      if (*e == '}' && e > s + 2) {
        size_t klen = (size_t)(e - (s + 2));
        for (int i = 0; i < nkeys; i++)
          if (strlen(keys[i]) == klen && !strncmp(keys[i], s + 2, klen)) { puts_(out, vals[i]); break; }
        s = e + 1;
        continue;
      }
    }
    put(out, s, 1);
    s++;
  }
}

/* One unit function (CsxJit::DoSpmvFnHook, CsxJit.hpp:413-545).  Returns the function name in fn. */
static int gen_unit(struct buf *defs, long pattern_id, int symmetric, char *fn, size_t fnlen) {
  long type = pattern_id / 10000, delta = pattern_id % 10000;
  char a[32], b[32];
  const char *tname = NULL, *keys[2], *vals[2];
  int nk = 1;
  if (type == 0) {
    tname = symmetric ? "delta_sym_tmpl.c" : "delta_tmpl.c";
    snprintf(a, sizeof a, "%ld", delta); keys[0] = "bits"; vals[0] = a;
    snprintf(fn, fnlen, "delta%ld_case", delta);
  } else if (type >= 1 && type <= 4) {
    static const char *base[] = {"horiz", "vert", "diag", "rdiag"};
    char t[64];
    snprintf(t, sizeof t, "%s%s_tmpl.c", base[type - 1], symmetric ? "_sym" : "");
    tname = strdup(t);
    snprintf(a, sizeof a, "%ld", delta); keys[0] = "delta"; vals[0] = a;
    snprintf(fn, fnlen, "%s%ld_case", base[type - 1], delta);
  } else if (type >= 5 && type <= 12) {
    long r = type - 4, c = delta;
    tname = symmetric ? "block_row_sym_tmpl.c" : (r == 1 ? "block_row_one_tmpl.c" : "block_row_tmpl.c");
    snprintf(a, sizeof a, "%ld", r); snprintf(b, sizeof b, "%ld", c);
    keys[0] = "r"; vals[0] = a; keys[1] = "c"; vals[1] = b; nk = 2;
    snprintf(fn, fnlen, "block_row_%ldx%ld_case", r, c);
  } else if (type >= 13 && type <= 20) {
    long r = delta, c = type - 12;
    tname = symmetric ? "block_col_sym_tmpl.c" : (c == 1 ? "block_col_one_tmpl.c" : "block_col_tmpl.c");
    snprintf(a, sizeof a, "%ld", r); snprintf(b, sizeof b, "%ld", c);
    keys[0] = "r"; vals[0] = a; keys[1] = "c"; vals[1] = b; nk = 2;
    snprintf(fn, fnlen, "block_col_%ldx%ld_case", r, c);
  } else {
    return -1;
  }
  const char *text = find_template(tname);
  if (!text) return -2;
  substitute(defs, text, keys, vals, nk);
  puts_(defs, "\n");
  return 0;
}

/* Writes the translation unit for one partition.  id_map is terminated by -1. */
int csxref_generate(const long *id_map, int row_jumps, int full_colind, int symmetric, const char *out_path) {
  struct buf defs = {0, 0, 0}, body = {0, 0, 0}, src = {0, 0, 0};
  puts_(&defs, "");
  puts_(&body, "");
  int n = 0;
  while (id_map[n] != -1) n++;
  char fn[64][64];
  for (int i = 0; i < n; i++)
    if (gen_unit(&defs, id_map[i], symmetric, fn[i], sizeof fn[i])) return -1;
  /* body hook, CsxJit.hpp:636-672: no switch for a single unit kind */
  const char *args = symmetric ? "(&ctl, size, &v, x, y, cur, &x_indx, &y_indx, scale_f);"
                               : "(&ctl, size, &v, &x_curr, &y_curr, scale_f);";
  char line[256];
  if (n == 1) {
    snprintf(line, sizeof line, "yr += %s%s", fn[0], args);
    puts_(&body, line);
  } else {
    puts_(&body, "switch (patt_id) {\n");
    for (int i = 0; i < n; i++) {
      snprintf(line, sizeof line, "\t\tcase %d:\n\t\t\tyr += %s%s\n\t\t\tbreak;\n", i, fn[i], args);
      puts_(&body, line);
    }
    puts_(&body, "\t\tdefault:\n\t\t\tfprintf(stderr, \"[BUG] unknown pattern\\n\");\n\t\t\texit(1);\n\t\t};");
  }
  /* new-row and next-x hooks, CsxJit.hpp:359-411 */
  const char *new_row, *next_x;
  if (!symmetric) {
    new_row = row_jumps ? "if (test_bit(&flags, CTL_RJMP_BIT))\n\t\t\t\ty_curr += ul_get(&ctl);\n\t\t\telse\n\t\t\t\ty_curr++;"
                        : "y_curr++;";
    next_x = full_colind ? "x_curr = x + u32_get(&ctl);" : "x_curr += ul_get(&ctl);";
  } else {
    new_row = row_jumps
                  ? "if (test_bit(&flags, CTL_RJMP_BIT)) {\n\t\t\t\tint jmp = ul_get(&ctl);\n\t\t\t\tfor (i = 0; i < jmp; i++) {\n"
                    "\t\t\t\t\ty[y_indx] += x[y_indx] * (*dv) * scale_f;\n\t\t\t\t\ty_indx++;\n\t\t\t\t\tdv++;\n\t\t\t\t}\n"
                    "\t\t\t} else {\n\t\t\t\ty[y_indx] += x[y_indx] * (*dv) * scale_f;\n\t\t\t\ty_indx++;\n\t\t\t\tdv++;\n\t\t\t}\n"
                  : "y[y_indx] += x[y_indx] * (*dv) * scale_f;\n\t\t\ty_indx++;\n\t\t\tdv++;\n";
    next_x = full_colind ? "x_indx = u32_get(&ctl);" : "x_indx += ul_get(&ctl);";
  }
  const char *main_t = find_template(symmetric ? "csx_sym_spmv_tmpl.c" : "csx_spmv_tmpl.c");
  if (!main_t) return -2;
  const char *keys[4] = {"spmv_func_definitions", "new_row_hook", "next_x", "body_hook"};
  const char *vals[4] = {defs.p, new_row, next_x, body.p};
  substitute(&src, main_t, keys, vals, 4);
  FILE *f = fopen(out_path, "w");
  if (!f) return -3;
  fwrite(src.p, 1, src.n, f);
  fclose(f);
  free(defs.p); free(body.p); free(src.p);
  return 0;
}

/* Writes the headers the templates include into <dir>/sparsex/... */
int csxref_write_headers(const char *dir) {
  char path[1024];
  snprintf(path, sizeof path, "%s/sparsex", dir); mkdir(path, 0755);
  snprintf(path, sizeof path, "%s/sparsex/internals", dir); mkdir(path, 0755);
  for (int i = 0; csxref_headers[i].name; i++) {
    snprintf(path, sizeof path, "%s/%s", dir, csxref_headers[i].name);
    FILE *f = fopen(path, "w");
    if (!f) return -1;
    fputs(csxref_headers[i].text, f);
    fclose(f);
  }
  return 0;
}
