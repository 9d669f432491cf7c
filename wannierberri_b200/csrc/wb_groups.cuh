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
};

// For every band n: g1[n], g2[n] = [ib1, ib2) of the kept group (or of the Fermi-sea group) that
// contains it, or g1[n] = -1.  label[ib1] = group label energy (mean of E, or -inf for the sea
// group), label = +inf for slots where no group starts.  Serial, O(nw); E ascending.
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
