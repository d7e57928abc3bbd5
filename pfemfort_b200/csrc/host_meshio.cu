// host_meshio.cu -- host-side mesh input for the drivers: the reference's text tables, parsed fast, and a binary
// container of the same arrays.  No CUDA in this file (built with nvcc only to share the build recipe).
//
// The reference reads every input file twice with list-directed READs (count pass + read pass,
// tetrapoissonparallelimpl1.F:216-355): at 48 M elements that is the largest part of its time to solution
// (SURVEY.md section 8f, rank 2).  The text format stays the compatibility path; `.pfemb` holds exactly the arrays
// the drivers build from the three (four) text files, so that a mesh is parsed once and then memory-mapped in.
//
// .pfemb layout (little endian):  char magic[8] = "PFEMB1\0\0";  int64 ndim, npElem, nNode, nElem, nDBC, nFBC;
//   double coords[ndim][nNode]     (the drivers' column-major coords(nNode, ndim), OLD numbering)
//   int32  conn[npElem][nElem]     (elemNodeConn(nElem, npElem), 1-based)            -- padded to 8 bytes
//   int32  dbc_node[nDBC], dbc_dof[nDBC] (each padded to 8 bytes);  double dbc_val[nDBC]
//   int32  fbc_node[nFBC], fbc_dof[nFBC] (each padded to 8 bytes);  double fbc_val[nFBC]
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "internal.cuh"

#define PFEM_EXPORT extern "C" __attribute__((visibility("default")))

static const char PFEMB_MAGIC[8] = {'P', 'F', 'E', 'M', 'B', '1', 0, 0};

// One number at *pp (after optional blanks).  Fast path: sign, up to 15 significant decimal digits, optional fraction --
// an exactly representable integer divided by an exactly representable power of ten, hence correctly rounded (the same
// double strtod returns); anything else (exponents, long mantissas, inf/nan) goes to strtod.  false: no number here.
static inline bool parse_number(const char *&p, double &v)
{
    static const double P10[16] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};
    const char *q = p;
    while (*q == ' ' || *q == '\t' || *q == '\r' || *q == ',') q++;
    const char *start = q;
    if (*q == '\n' || *q == 0) return false;            // end of the row (strtod would run on into the next line)
    bool neg = false;
    if (*q == '-' || *q == '+') { neg = *q == '-'; q++; }
    uint64_t mant = 0;
    int digits = 0, frac = 0;
    while (*q >= '0' && *q <= '9') { mant = mant * 10 + (uint64_t)(*q - '0'); digits++; q++; }
    if (*q == '.') {
        q++;
        while (*q >= '0' && *q <= '9') { mant = mant * 10 + (uint64_t)(*q - '0'); digits++; frac++; q++; }
    }
    const bool simple = digits > 0 && digits <= 15 && *q != 'e' && *q != 'E' && *q != 'd' && *q != 'D';
    if (simple) {
        v = (double)mant / P10[frac];
        if (neg) v = -v;
        p = q;
        return true;
    }
    // slow path on a private copy of the token, with Fortran's D exponent letter mapped to E
    char tok[64];
    int n = 0;
    while (n < 63 && start[n] != ' ' && start[n] != '\t' && start[n] != '\r' && start[n] != ',' && start[n] != '\n' && start[n] != 0) {
        tok[n] = (start[n] == 'd' || start[n] == 'D') ? 'e' : start[n];
        n++;
    }
    tok[n] = 0;
    char *next = nullptr;
    v = strtod(tok, &next);
    if (next == tok) return false;
    p = start + (next - tok);
    return true;
}

// Read a whitespace-separated numeric table (the reference's `id v1 v2 ...` rows).  Rows with fewer than ncols numbers
// are skipped (blank / trailing lines), extra numbers on a row are ignored, like the drivers' READ.
// Two-call protocol: out == NULL returns the number of rows (a tokenising scan, no conversions); otherwise fills
// out[c*nrows_cap + r] (column-major, as the Fortran arrays) for at most nrows_cap rows.  Returns -1 if the file cannot
// be read.  (The reference makes the same two passes with list-directed READs, tetrapoissonparallelimpl1.F:216-238.)
PFEM_EXPORT long long pfem_host_read_table(const char *path, int ncols, double *out, long long nrows_cap)
{
    FILE *f = fopen(path, "rb");
    if (!f) { pfem::set_error("File ... %s does not exist", path); return -1; }
    fseek(f, 0, SEEK_END);
    const long long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)size + 2);
    if (size > 0 && fread(buf.data(), 1, (size_t)size, f) != (size_t)size) { fclose(f); pfem::set_error("cannot read %s", path); return -1; }
    fclose(f);
    buf[size] = '\n';
    buf[size + 1] = 0;
    long long nrows = 0;
    const char *p = buf.data(), *end = buf.data() + size;
    if (!out) {                                        // count pass: a row counts when it holds at least ncols tokens
        while (p < end) {
            int tokens = 0;
            while (*p != '\n') {
                while (*p == ' ' || *p == '\t' || *p == '\r' || *p == ',') p++;
                if (*p == '\n') break;
                tokens++;
                while (*p != ' ' && *p != '\t' && *p != '\r' && *p != ',' && *p != '\n') p++;
            }
            nrows += tokens >= ncols;
            p++;
        }
        return nrows;
    }
    std::vector<double> row(ncols);
    while (p < end) {
        int got = 0;
        while (got < ncols && parse_number(p, row[got])) got++;
        while (*p != '\n') p++;                        // extra columns are ignored
        p++;
        if (got == ncols) {
            if (nrows < nrows_cap)
                for (int c = 0; c < ncols; c++) out[(size_t)c * nrows_cap + nrows] = row[c];
            nrows++;
        }
    }
    return nrows;
}

static bool put(FILE *f, const void *p, size_t bytes) { return bytes == 0 || fwrite(p, 1, bytes, f) == bytes; }
static bool put_i32(FILE *f, const int *p, long long n)
{
    static const char pad[4] = {0, 0, 0, 0};
    return put(f, p, (size_t)n * 4) && ((n & 1) == 0 || put(f, pad, 4));
}

PFEM_EXPORT int pfem_host_mesh_write_binary(const char *path, int ndim, int npElem, int nNode, int nElem, const double *coords,
                                            const int *conn, int nDBC, const int *dbc_node, const int *dbc_dof,
                                            const double *dbc_val, int nFBC, const int *fbc_node, const int *fbc_dof,
                                            const double *fbc_val)
{
    FILE *f = fopen(path, "wb");
    if (!f) { pfem::set_error("cannot create %s: %s", path, strerror(errno)); return PFEM_ERR_ARG; }
    const int64_t hdr[6] = {ndim, npElem, nNode, nElem, nDBC, nFBC};
    bool ok = put(f, PFEMB_MAGIC, 8) && put(f, hdr, sizeof hdr) && put(f, coords, (size_t)ndim * nNode * 8) &&
              put_i32(f, conn, (long long)npElem * nElem) && put_i32(f, dbc_node, nDBC) && put_i32(f, dbc_dof, nDBC) &&
              put(f, dbc_val, (size_t)nDBC * 8) && put_i32(f, fbc_node, nFBC) && put_i32(f, fbc_dof, nFBC) &&
              put(f, fbc_val, (size_t)nFBC * 8);
    ok = fclose(f) == 0 && ok;
    if (!ok) { pfem::set_error("short write to %s", path); return PFEM_ERR_ARG; }
    return PFEM_OK;
}

// sizes[6] = ndim, npElem, nNode, nElem, nDBC, nFBC.  Returns PFEM_ERR_ARG if the file is not a .pfemb container.
PFEM_EXPORT int pfem_host_mesh_read_binary_header(const char *path, long long sizes[6])
{
    FILE *f = fopen(path, "rb");
    if (!f) { pfem::set_error("File ... %s does not exist", path); return PFEM_ERR_ARG; }
    char magic[8];
    int64_t hdr[6];
    const bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, PFEMB_MAGIC, 8) == 0 && fread(hdr, 8, 6, f) == 6;
    fclose(f);
    if (!ok) { pfem::set_error("%s is not a PFEMB1 mesh container", path); return PFEM_ERR_ARG; }
    for (int i = 0; i < 6; i++) sizes[i] = hdr[i];
    if (hdr[0] < 2 || hdr[0] > 3 || hdr[1] < 3 || hdr[1] > 4 || hdr[2] <= 0 || hdr[3] <= 0 || hdr[4] < 0 || hdr[5] < 0 ||
        hdr[2] >= (1LL << 31) || hdr[3] >= (1LL << 31)) {
        pfem::set_error("%s: implausible header", path);
        return PFEM_ERR_ARG;
    }
    return PFEM_OK;
}

static bool get(FILE *f, void *p, size_t bytes) { return bytes == 0 || fread(p, 1, bytes, f) == bytes; }
static bool get_i32(FILE *f, int *p, long long n)
{
    char pad[4];
    return get(f, p, (size_t)n * 4) && ((n & 1) == 0 || get(f, pad, 4));
}

// Arrays sized from the header by the caller (any pointer of an empty section may be NULL).
PFEM_EXPORT int pfem_host_mesh_read_binary(const char *path, double *coords, int *conn, int *dbc_node, int *dbc_dof,
                                           double *dbc_val, int *fbc_node, int *fbc_dof, double *fbc_val)
{
    long long s[6];
    PFEM_TRY(pfem_host_mesh_read_binary_header(path, s));
    FILE *f = fopen(path, "rb");
    if (!f) { pfem::set_error("File ... %s does not exist", path); return PFEM_ERR_ARG; }
    fseek(f, 8 + 6 * 8, SEEK_SET);
    const bool ok = get(f, coords, (size_t)(s[0] * s[2]) * 8) && get_i32(f, conn, s[1] * s[3]) && get_i32(f, dbc_node, s[4]) &&
                    get_i32(f, dbc_dof, s[4]) && get(f, dbc_val, (size_t)s[4] * 8) && get_i32(f, fbc_node, s[5]) &&
                    get_i32(f, fbc_dof, s[5]) && get(f, fbc_val, (size_t)s[5] * 8);
    fclose(f);
    if (!ok) { pfem::set_error("%s: truncated", path); return PFEM_ERR_ARG; }
    return PFEM_OK;
}
