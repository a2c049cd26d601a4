/* A plain-C host of the drop-in boundary: loads a pattern IR (written by the Python front end), builds the model
 * on the current GPU through include/exa_b200.h and evaluates hess_coord! with HOST buffers.  No Python, no torch.
 * usage: c_host_example <ir file> <x file> <y file> <obj_weight>   -> prints nnzh and two checksums */
#include <stdio.h>
#include <stdlib.h>
#include "exa_b200.h"

static void* slurp(const char* path, size_t* n) {
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  fseek(f, 0, SEEK_END); *n = (size_t)ftell(f); fseek(f, 0, SEEK_SET);
  void* p = malloc(*n ? *n : 1);
  if (fread(p, 1, *n, f) != *n) { perror("read"); exit(2); }
  fclose(f);
  return p;
}

int main(int argc, char** argv) {
  if (argc < 5) { fprintf(stderr, "usage: %s ir x y obj_weight\n", argv[0]); return 2; }
  size_t nir, nx, ny;
  void* ir = slurp(argv[1], &nir);
  double* x = (double*)slurp(argv[2], &nx);
  double* y = (double*)slurp(argv[3], &ny);
  exb_options opt = {-1, 0, 1, 0, 0, NULL};
  exb_model* m = NULL;
  if (exb_create(ir, nir, NULL, 0, &opt, &m)) { fprintf(stderr, "exb_create: %s\n", exb_last_error()); return 1; }
  int64_t d[EXB_NDIMS];
  exb_dims(m, d);
  if ((size_t)d[0] * 8 != nx || (size_t)d[1] * 8 != ny) { fprintf(stderr, "size mismatch\n"); return 1; }
  double* h = (double*)malloc((size_t)(d[3] ? d[3] : 1) * 8);
  if (exb_host_hess(m, x, y, atof(argv[4]), h)) { fprintf(stderr, "exb_host_hess: %s\n", exb_last_error()); return 1; }
  double s = 0, w = 0, obj = 0;
  for (int64_t k = 0; k < d[3]; k++) { s += h[k]; w += h[k] * (double)((k % 97) + 1); }
  if (exb_host_obj(m, x, &obj)) { fprintf(stderr, "exb_host_obj: %s\n", exb_last_error()); return 1; }
  printf("%lld %.17g %.17g %.17g\n", (long long)d[3], s, w, obj);
  exb_destroy(m);
  return 0;
}
