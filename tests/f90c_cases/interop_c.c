/* C side of tests/f90c_cases/interop.F90 (test infrastructure for oracle/f90c.py) */
#include <stddef.h>
typedef struct { double scale; int n; void *data; void *spare; } cblock;
int c_scale(cblock *b, double factor) {
  double *d = (double *)b->data;
  for (int i = 0; i < b->n; ++i) d[i] *= factor * b->scale;
  b->spare = b->data ? b->data : (void *)b;    /* something non-NULL */
  return b->n;
}
const char *c_name(void) { return "thirteen char"; }
int c_sum3(const int *v) { return v[0] + v[1] + v[2]; }
