/* TEST DOUBLE of the part of the C ABI (include/pfem_b200.h) that the edited reference PROGRAM drives: same symbols and
 * prototypes, no GPU.  It stores what crosses the boundary so that tests/test_refrun_dropin.py can check, on the CPU, the
 * marshalling of include/pfem_b200.f90 + the INTEGRATION.md diff executed by oracle/refrun (the real library needs a B200;
 * the same program runs against it in tests/test_gpu_zzzz_reference_vectors.py).  Never shipped, never linked by the product. */
#include <stdlib.h>
#include <string.h>

typedef struct {
    int device, rank, nranks, size_local, size_global, kind, nElem, nNode, nsize, n_applied, calls[16], ncalls;
    int n_added, added_row[64]; double added_val[64];
    int *diag_nnz, *offdiag_nnz, *conn, *old, *edof;
    double *coords, *applied, elemData[8], timeData[8];
} fake_t;

static fake_t *last;
static void note(fake_t *f, int id) { if (f->ncalls < 16) f->calls[f->ncalls++] = id; }
static int *dupi(const int *p, size_t n) { int *q = malloc(n * sizeof(int)); memcpy(q, p, n * sizeof(int)); return q; }
static double *dupd(const double *p, size_t n) { double *q = malloc(n * sizeof(double)); memcpy(q, p, n * sizeof(double)); return q; }

int pfem_solver_create(void **h, int device, int rank, int nranks, const void *id) {
    fake_t *f = calloc(1, sizeof(fake_t));
    f->device = device; f->rank = rank; f->nranks = nranks; (void)id;
    *h = f; last = f; note(f, 1); return 0;
}
int pfem_solver_initialise(fake_t *f, int sl, int sg, const int *d, const int *o) {
    f->size_local = sl; f->size_global = sg; f->diag_nnz = dupi(d, sl); f->offdiag_nnz = dupi(o, sl); note(f, 2); return 0;
}
int pfem_solver_set_mesh(fake_t *f, int kind, int nElem, const int *conn, int nNode, const double *coords, const int *old) {
    int npe = (kind == 0 || kind == 2) ? 3 : 4, ndim = (kind == 0 || kind == 2) ? 2 : 3;
    f->kind = kind; f->nElem = nElem; f->nNode = nNode;
    f->conn = dupi(conn, (size_t)npe * nElem); f->coords = dupd(coords, (size_t)ndim * nNode); f->old = old ? dupi(old, nNode) : NULL;
    note(f, 3); return 0;
}
int pfem_solver_set_pattern(fake_t *f, int nElem, int nsize, const int *edof) {
    if (nElem != f->nElem) return 2;
    f->nsize = nsize; f->edof = dupi(edof, (size_t)nsize * nElem); note(f, 4); return 0;
}
int pfem_solver_set_zero(fake_t *f) { note(f, 5); return 0; }
int pfem_solver_set_applied(fake_t *f, const double *a, int n) { f->n_applied = n; f->applied = dupd(a, n); note(f, 6); return 0; }
int pfem_solver_assemble(fake_t *f, const double *ed, const double *td, int *nneg) {
    memcpy(f->elemData, ed, sizeof f->elemData); memcpy(f->timeData, td, sizeof f->timeData); *nneg = 0; note(f, 7); return 0;
}
int pfem_solver_add_value(fake_t *f, int row, double val) {
    if (f->n_added < 64) { f->added_row[f->n_added] = row; f->added_val[f->n_added] = val; }
    f->n_added++; return 0;
}
int pfem_solver_factorise_and_solve(fake_t *f) { note(f, 8); return 0; }
int pfem_solver_get_solution(fake_t *f, double *x) { for (int i = 0; i < f->size_global; i++) x[i] = 1000.0 + i; note(f, 9); return 0; }
int pfem_solver_free(fake_t *f) { note(f, 10); return 0; }   /* kept alive for the test to read */
fake_t *fake_last(void) { return last; }
int fake_scalar(fake_t *f, int which) {
    int v[] = {f->device, f->rank, f->nranks, f->size_local, f->size_global, f->kind, f->nElem, f->nNode, f->nsize, f->n_applied, f->ncalls,
               f->n_added};
    return v[which];
}
const int *fake_ints(fake_t *f, int which) { const int *p[] = {f->diag_nnz, f->offdiag_nnz, f->conn, f->old, f->edof, f->calls, f->added_row}; return p[which]; }
const double *fake_doubles(fake_t *f, int which) { const double *p[] = {f->coords, f->applied, f->elemData, f->timeData, f->added_val}; return p[which]; }
