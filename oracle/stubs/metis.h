/*
 * Stand-in for <metis.h>.  TEST INFRASTRUCTURE ONLY.
 * The one-rank oracle build never partitions (TACSCreator skips METIS when
 * size == 1 or a partition array is supplied) and uses NATURAL_ORDER, so the
 * three entry points only need to exist; reaching one is an error.
 */
#ifndef A2DS_ORACLE_STUB_METIS_H
#define A2DS_ORACLE_STUB_METIS_H
#include <stdio.h>
#include <stdlib.h>
typedef int idx_t;
typedef float real_t;
#define METIS_NOPTIONS 40
#define METIS_OPTION_NUMBERING 17
#define METIS_OK 1
static inline int METIS_SetDefaultOptions(idx_t *o) {
  for (int i = 0; i < METIS_NOPTIONS; i++) o[i] = -1;
  return METIS_OK;
}
static inline int a2ds_metis_die(const char *w) {
  fprintf(stderr, "[oracle metis stub] %s is not available\n", w);
  abort();
  return 0;
}
static inline int METIS_PartGraphRecursive(idx_t *a, idx_t *b, idx_t *c, idx_t *d, idx_t *e, idx_t *f, idx_t *g, idx_t *h, real_t *i, real_t *j, idx_t *k, idx_t *l, idx_t *m) {
  (void)a;(void)b;(void)c;(void)d;(void)e;(void)f;(void)g;(void)h;(void)i;(void)j;(void)k;(void)l;(void)m;
  return a2ds_metis_die("METIS_PartGraphRecursive");
}
static inline int METIS_PartGraphKway(idx_t *a, idx_t *b, idx_t *c, idx_t *d, idx_t *e, idx_t *f, idx_t *g, idx_t *h, real_t *i, real_t *j, idx_t *k, idx_t *l, idx_t *m) {
  (void)a;(void)b;(void)c;(void)d;(void)e;(void)f;(void)g;(void)h;(void)i;(void)j;(void)k;(void)l;(void)m;
  return a2ds_metis_die("METIS_PartGraphKway");
}
static inline int METIS_NodeND(idx_t *a, idx_t *b, idx_t *c, idx_t *d, idx_t *e, idx_t *f, idx_t *g) {
  (void)a;(void)b;(void)c;(void)d;(void)e;(void)f;(void)g;
  return a2ds_metis_die("METIS_NodeND");
}
#endif
