// Band groups at one k-point -- decision-exact restatement of
//   get_borders / get_bands_in_range / get_bands_below_range   (grid/tetrahedron.py:132-162)
//   Data_K.get_bands_in_range_groups_ik                        (data_K/data_K.py:172-186)
#pragma once
#include "wb_common.cuh"

struct WbWindow {
    double EFmin, EFmax, dEF;  // static.py:54-57 (already widened by extraEf * dEF)
    double degen_thresh;
    int degen_Kramers;
    int sea;                   // fder == 0
    int nEFx;                  // nEF_extra
    // tetrahedron method (grid/tetrahedron.py:140-162 with Ebandmin / Ebandmax): per-band minimum / maximum of the
    // energy over the k-point's cell (centre + 8 corners), [nk][nw]; nullptr = plain rule on the centre energies
    const double* Ebmin;
    const double* Ebmax;
    // tetrahedron method: hole_like (der = -1: no Fermi-sea group, but the group of the bands ABOVE EFmax, up to Emax_holes),
    // lower edge of the Fermi-sea group (tetrahedron.py:246-266)
    int holes;
    double Emin_sea, Emax_holes;
};

// For every band n: g1[n], g2[n] = [ib1, ib2) of the kept group (or of the Fermi-sea group) that
// contains it, or g1[n] = -1.  label[ib1] = group label energy (mean of E, or -inf for the sea
// group), label = +inf for slots where no group starts.  Serial, O(nw); E ascending.
// tetrahedron variant: emin / emax = this k-point's rows of WbWindow::Ebmin / Ebmax.
__device__ inline void wb_band_groups_tetra(const double* E, const double* __restrict__ emin, const double* __restrict__ emax,
                                            int nw, const WbWindow& w, short* g1, short* g2, double* label) {
    for (int n = 0; n < nw; n++) { g1[n] = -1; g2[n] = -1; label[n] = CUDART_INF; }
    int prev = -1, first_kept = -1;
    for (int pos = 0; pos <= nw; pos++) {
        bool border = (pos == 0) || (pos == nw) || (E[pos] - E[pos - 1] > w.degen_thresh);
        if (w.degen_Kramers && (pos & 1)) border = false;
        if (!border) continue;
        if (prev >= 0) {
            const int a = prev, b = pos;
            double mx = emax[a], mn = emin[a], s = E[a];
            for (int n = a + 1; n < b; n++) { mx = fmax(mx, emax[n]); mn = fmin(mn, emin[n]); s += E[n]; }
            if (mx >= w.EFmin && mn <= w.EFmax) {   // get_bands_in_range
                label[a] = s / (double)(b - a);
                for (int n = a; n < b; n++) { g1[n] = (short)a; g2[n] = (short)b; }
                if (first_kept < 0) first_kept = a;
            }
        }
        prev = pos;
    }
    if (w.sea && !w.holes) {   // get_bands_below_range(eFermi[0] | Emin, Ebandmax = Emax)  (tetrahedron.py:246-256)
        int bandmax = 0, bandmin = 0;
        for (int n = 0; n < nw; n++) {
            if (emax[n] < w.EFmin) bandmax = n + 1;
            if (emax[n] < w.Emin_sea) bandmin = n + 1;
        }
        if (first_kept >= 0) bandmax = min(bandmax, first_kept);
        if (bandmax > bandmin) {
            label[bandmin] = -CUDART_INF;
            for (int n = bandmin; n < bandmax; n++) { g1[n] = (short)bandmin; g2[n] = (short)bandmax; }
        }
    }
    if (w.holes) {   // get_bands_above_range(eFermi[-1] | Emax, Ebandmin = Emin)  (tetrahedron.py:258-266)
        int bandmin = nw, bandmax = nw, last_end = -1;
        for (int n = nw - 1; n >= 0; n--) {
            if (emin[n] > w.EFmax) bandmin = n;
            if (emin[n] > w.Emax_holes) bandmax = n;
        }
        for (int n = 0; n < nw; n++)
            if (g1[n] >= 0) last_end = g2[n];
        if (last_end >= 0) bandmin = max(bandmin, last_end);
        if (bandmax > bandmin) {
            label[bandmin] = -CUDART_INF;
            for (int n = bandmin; n < bandmax; n++) { g1[n] = (short)bandmin; g2[n] = (short)bandmax; }
        }
    }
}

__device__ inline void wb_band_groups(const double* E, int nw, const WbWindow& w, short* g1, short* g2,
                                      double* label) {
    for (int n = 0; n < nw; n++) { g1[n] = -1; g2[n] = -1; label[n] = CUDART_INF; }
    int prev = -1;
    int first_kept = -1;
    for (int pos = 0; pos <= nw; pos++) {
        bool border = (pos == 0) || (pos == nw) || (E[pos] - E[pos - 1] > w.degen_thresh);
        if (w.degen_Kramers && (pos & 1)) border = false;
        if (!border) continue;
        if (prev >= 0) {
            int a = prev, b = pos;
            if (E[b - 1] >= w.EFmin && E[a] <= w.EFmax) {
                double s = E[a];
                for (int n = a + 1; n < b; n++) s += E[n];
                label[a] = s / (double)(b - a);
                for (int n = a; n < b; n++) { g1[n] = (short)a; g2[n] = (short)b; }
                if (first_kept < 0) first_kept = a;
            }
        }
        prev = pos;
    }
    if (w.sea) {
        int bandmax = 0;
        for (int n = 0; n < nw; n++)
            if (E[n] < w.EFmin) bandmax = n + 1;
        if (first_kept >= 0) bandmax = min(bandmax, first_kept);
        if (bandmax > 0) {
            label[0] = -CUDART_INF;
            for (int n = 0; n < bandmax; n++) { g1[n] = 0; g2[n] = (short)bandmax; }
        }
    }
}

// Same decisions, evaluated by one full warp for nw <= 32 (lane n = band n; E in shared memory):
// borders -> 64-bit ballot mask, group start / end by bit scans.  Keeps the per-k-point band-group
// logic off the critical path of the rotation kernels (the serial version costs ~nw dependent
// shared-memory round trips on a single thread while the rest of the CTA waits at a barrier).
__device__ inline void wb_band_groups_warp(const double* Es, int nw, const WbWindow& w, short* g1, short* g2,
                                           double* label, int lane) {
    const unsigned full = 0xffffffffu;
    double En = (lane < nw) ? Es[lane] : CUDART_INF;
    bool brd = false;
    if (lane < nw) {
        brd = (lane == 0) || (En - Es[lane - 1] > w.degen_thresh);
        if (w.degen_Kramers && (lane & 1)) brd = false;
    }
    unsigned long long B = __ballot_sync(full, brd);
    if (!(w.degen_Kramers && (nw & 1))) B |= (1ull << nw);
    int a = -1, b = -1;
    bool kept = false;
    if (lane < nw) {
        unsigned long long lo = B & ((2ull << lane) - 1ull);  // borders at positions <= lane (bit 0 is always set)
        a = 63 - __clzll((long long)lo);
        unsigned long long hi = B >> (lane + 1);
        if (hi) {
            b = lane + __ffsll((long long)hi);
            kept = (Es[b - 1] >= w.EFmin) && (Es[a] <= w.EFmax);
        }
    }
    unsigned keptstart = __ballot_sync(full, kept && lane == a);
    int first_kept = keptstart ? (__ffs(keptstart) - 1) : -1;
    int G1 = kept ? a : -1, G2 = kept ? b : -1;
    double lab = CUDART_INF;
    if (kept && lane == a) {
        double s = Es[a];
        for (int n = a + 1; n < b; n++) s += Es[n];
        lab = s / (double)(b - a);
    }
    if (w.sea) {
        unsigned below = __ballot_sync(full, lane < nw && En < w.EFmin);
        int bandmax = below ? (32 - __clz(below)) : 0;
        if (first_kept >= 0) bandmax = min(bandmax, first_kept);
        if (bandmax > 0) {
            if (lane < bandmax) { G1 = 0; G2 = bandmax; }
            if (lane == 0) lab = -CUDART_INF;
        }
    }
    if (lane < nw) { g1[lane] = (short)G1; g2[lane] = (short)G2; label[lane] = lab; }
}
