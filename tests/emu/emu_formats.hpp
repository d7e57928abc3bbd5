// emu_formats.hpp -- TEST INFRASTRUCTURE ONLY: host restatement of the device-side array formats the pattern pass
// produces (pattern.cu: pack_conn/pack_dof, pack_xyz, conn4, the row incidence lists and the row-gather incidence
// streams with their slot bytes), from the driver-level arrays.  Shared by the kernel emulators.
#pragma once
#include <algorithm>
#include <vector>

struct EmuFormats {
    int npe = 0, ndof = 0, nsize = 0, ndim = 0, rec_ints = 0, stride = 0, words = 0;
    std::vector<int> erec, conn4, rinc_ptr, rinc, ainc;
    std::vector<long long> ainc_off;
    std::vector<double> xyz;
};

// kind: 0 Poisson tria, 1 Poisson tet, 2 elasticity tria, 3 elasticity tet.
// conn0: [npe][nElem] 0-based NEW node ids; edof: [nsize][nElem] global dof ids (-1 Dirichlet); xyz_soa: [ndim][nNode] NEW
// numbering; rowptr/col: CSR pattern of the owned rows [row_lo, row_lo+nloc) with global columns.
inline void emu_build_formats(int kind, int nElem, int nNode, const int *conn0, const int *edof, const double *xyz_soa,
                              int row_lo, int nloc, const int *rowptr, const int *col, EmuFormats &f)
{
    f.npe = (kind == 0 || kind == 2) ? 3 : 4;
    f.ndof = kind == 0 || kind == 1 ? 1 : (kind == 2 ? 2 : 3);
    f.ndim = (kind == 0 || kind == 2) ? 2 : 3;
    f.nsize = f.npe * f.ndof;
    f.rec_ints = ((f.npe + f.nsize + 3) / 4) * 4;
    f.stride = f.ndim == 3 ? 4 : 2;
    f.words = f.nsize <= 4 ? 2 : 4;
    const int npe = f.npe, nsize = f.nsize;
    f.erec.assign((size_t)nElem * f.rec_ints, -1);
    f.conn4.resize((size_t)nElem * 4);
    for (int e = 0; e < nElem; e++) {
        for (int i = 0; i < npe; i++) f.erec[(size_t)e * f.rec_ints + i] = conn0[(size_t)i * nElem + e];
        for (int k = 0; k < nsize; k++) f.erec[(size_t)e * f.rec_ints + npe + k] = edof[(size_t)k * nElem + e];
        for (int i = 0; i < 4; i++) f.conn4[(size_t)e * 4 + i] = conn0[(size_t)(i < npe ? i : npe - 1) * nElem + e];
    }
    f.xyz.assign((size_t)nNode * f.stride, 0.0);
    for (int n = 0; n < nNode; n++)
        for (int d = 0; d < f.ndim; d++) f.xyz[(size_t)n * f.stride + d] = xyz_soa[(size_t)d * nNode + n];
    f.rinc_ptr.assign(nloc + 1, 0);
    for (int e = 0; e < nElem; e++)
        for (int k = 0; k < nsize; k++) {
            const int d = edof[(size_t)k * nElem + e];
            if (d >= row_lo && d < row_lo + nloc) f.rinc_ptr[d - row_lo + 1]++;
        }
    for (int r = 0; r < nloc; r++) f.rinc_ptr[r + 1] += f.rinc_ptr[r];
    f.rinc.assign(f.rinc_ptr[nloc] > 0 ? f.rinc_ptr[nloc] : 1, 0);
    std::vector<int> cur(f.rinc_ptr.begin(), f.rinc_ptr.end() - 1);
    for (int e = 0; e < nElem; e++)
        for (int k = 0; k < nsize; k++) {
            const int d = edof[(size_t)k * nElem + e];
            if (d >= row_lo && d < row_lo + nloc) f.rinc[cur[d - row_lo]++] = e * nsize + k;
        }
    const int nslices = (nloc + 31) / 32;
    f.ainc_off.assign(nslices + 1, 0);
    for (int s = 0; s < nslices; s++) {
        int w = 0;
        for (int l = 0; l < 32 && s * 32 + l < nloc; l++) w = std::max(w, f.rinc_ptr[s * 32 + l + 1] - f.rinc_ptr[s * 32 + l]);
        f.ainc_off[s + 1] = f.ainc_off[s] + (long long)w * 32;
    }
    f.ainc.assign((size_t)f.ainc_off[nslices] * f.words + 4, -1);
    for (int r = 0; r < nloc; r++) {
        const int c0 = rowptr[r], len = rowptr[r + 1] - c0;
        int m = 0;
        for (int q = f.rinc_ptr[r]; q < f.rinc_ptr[r + 1]; q++, m++) {
            const int code = f.rinc[q], e = code / nsize;
            unsigned int w[3] = {0u, 0u, 0u};             // unused bytes 0x00, Dirichlet dofs 0xFF
            for (int j = 0; j < nsize; j++) {
                const int c = edof[(size_t)j * nElem + e];
                unsigned int sl = 255u;
                if (c >= 0) sl = (unsigned int)(std::lower_bound(col + c0, col + c0 + len, c) - (col + c0));
                w[j >> 2] |= (sl & 255u) << (8 * (j & 3));
            }
            int *out = f.ainc.data() + (size_t)(f.ainc_off[r >> 5] + (r & 31) + (long long)m * 32) * f.words;
            out[0] = code;
            for (int q2 = 1; q2 < f.words; q2++) out[q2] = (int)w[q2 - 1];
        }
    }
}
