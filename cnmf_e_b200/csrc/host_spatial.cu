// host_spatial.cu -- host-side helpers of the reference that sit right next to the hot path:
//   determine_search_location(A, 'ellipse', params)   ca_source_extraction/utilities/determine_search_location.m:57-92
//   post_process_spatial / connectivity_constraint    @Sources2D/post_process_spatial.m:19-32, endoscope/connectivity_constraint.m
//   graph_connected_comp                               utilities/graph_conn_comp_mex.cpp (the reference's only native file)
// They are host code in the reference too (MATLAB, per neuron on a small crop); here they are plain C++ on the CSC matrix so that
// the host mirror / MEX gateway need not leave the library between cnmfe_update_spatial calls.  No CUDA in this file.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>
#include "common.cuh"
#include "internal.h"

using cnmfe::set_error;

// connectivity_constraint(img, thr, sz): ai_open = imopen(img, strel('square', sz)); temp = ai_open > max(img)*thr;
// l = bwlabel(temp, 4); img(l ~= l(ind_max)) = 0  with ind_max = first arg-max of img (column-major).
// MATLAB's flat morphology ignores pixels outside the image (erosion pads +Inf, dilation -Inf).  Note the reference's
// behaviour when the arg-max pixel is NOT in the opened mask: l(ind_max) = 0, so every labelled component is removed and the
// unlabelled remainder is kept -- replicated.
extern "C" int cnmfe_connectivity_constraint(int d1, int d2, int K, const int64_t* jc, const int64_t* ir, double* pr,
                                             double thr, int sz) {
    if (d1 <= 0 || d2 <= 0 || K < 0 || !jc || (K > 0 && jc[K] > 0 && (!ir || !pr))) { set_error("cnmfe_connectivity_constraint: bad arguments"); return -1; }
    if (sz < 1 || (sz & 1) == 0) { set_error("cnmfe_connectivity_constraint: sz must be odd (strel('square', sz) centred)"); return -1; }
    const int h = sz / 2, margin = 2 * h;
    std::vector<double> img, ero, opn;
    std::vector<unsigned char> mask, keep;
    std::vector<int> stack;
    for (int k = 0; k < K; ++k) {
        const int64_t e0 = jc[k], e1 = jc[k + 1];
        // bounding box, maximum and its first position (entries are sorted by linear index = column-major order)
        int r0 = d1, r1 = -1, c0 = d2, c1 = -1;
        double vmax = 0.0;
        int64_t imax = 0;
        bool have = false, anynz = false;
        for (int64_t e = e0; e < e1; ++e) {
            const double v = pr[e];
            if (v == 0.0) continue;
            anynz = true;
            const int r = (int)(ir[e] % d1), c = (int)(ir[e] / d1);
            r0 = std::min(r0, r); r1 = std::max(r1, r); c0 = std::min(c0, c); c1 = std::max(c1, c);
            if (!have || v > vmax) { vmax = v; imax = ir[e]; have = true; }
        }
        if (!anynz) continue;
        // max(img(:)) over the FULL image: implicit zeros count when every stored value is negative
        if (vmax < 0.0 && (int64_t)d1 * d2 > (e1 - e0)) { vmax = 0.0; /* arg-max = first implicit zero */
            int64_t pos = 0, e = e0;
            while (e < e1 && ir[e] == pos && pr[e] != 0.0) { ++pos; ++e; }   // first linear index that is zero
            imax = pos;
        }
        const int R0 = std::max(0, r0 - margin), R1 = std::min(d1 - 1, r1 + margin);
        const int C0 = std::max(0, c0 - margin), C1 = std::min(d2 - 1, c1 + margin);
        const int nr = R1 - R0 + 1, nc = C1 - C0 + 1;
        img.assign((size_t)nr * nc, 0.0);
        for (int64_t e = e0; e < e1; ++e) {
            const int r = (int)(ir[e] % d1), c = (int)(ir[e] / d1);
            if (r >= R0 && r <= R1 && c >= C0 && c <= C1) img[(size_t)(c - C0) * nr + (r - R0)] = pr[e];
        }
        // pixels of the FOV outside the crop are zeros (the crop has a 2h margin); pixels outside the FOV do not count
        auto at = [&](const std::vector<double>& a, int r, int c, bool* inside) -> double {
            *inside = (r >= 0 && r < d1 && c >= 0 && c < d2);
            if (!*inside) return 0.0;
            if (r < R0 || r > R1 || c < C0 || c > C1) return 0.0;
            return a[(size_t)(c - C0) * nr + (r - R0)];
        };
        ero.assign(img.size(), 0.0);
        for (int c = C0; c <= C1; ++c)
            for (int r = R0; r <= R1; ++r) {
                double m = std::numeric_limits<double>::infinity();
                for (int dc = -h; dc <= h; ++dc)
                    for (int dr = -h; dr <= h; ++dr) {
                        bool in;
                        const double v = at(img, r + dr, c + dc, &in);
                        if (in) m = std::min(m, v);
                    }
                ero[(size_t)(c - C0) * nr + (r - R0)] = m;
            }
        opn.assign(img.size(), 0.0);
        for (int c = C0; c <= C1; ++c)
            for (int r = R0; r <= R1; ++r) {
                double m = -std::numeric_limits<double>::infinity();
                for (int dc = -h; dc <= h; ++dc)
                    for (int dr = -h; dr <= h; ++dr) {
                        bool in;
                        const double v = at(ero, r + dr, c + dc, &in);
                        if (in) m = std::max(m, v);
                    }
                opn[(size_t)(c - C0) * nr + (r - R0)] = m;
            }
        const double level = vmax * thr;
        mask.assign(img.size(), 0);
        for (size_t i = 0; i < img.size(); ++i) mask[i] = opn[i] > level ? 1 : 0;
        // component of the arg-max pixel (4-connected), or -- if that pixel is not in the mask -- the unlabelled remainder
        const int rm = (int)(imax % d1), cm = (int)(imax / d1);
        const bool max_in_crop = (rm >= R0 && rm <= R1 && cm >= C0 && cm <= C1);
        const bool seeded = max_in_crop && mask[(size_t)(cm - C0) * nr + (rm - R0)];
        keep.assign(img.size(), 0);
        if (seeded) {
            stack.clear();
            stack.push_back((cm - C0) * nr + (rm - R0));
            keep[stack.back()] = 1;
            while (!stack.empty()) {
                const int p = stack.back(); stack.pop_back();
                const int r = p % nr, c = p / nr;
                const int nb[4][2] = {{r - 1, c}, {r + 1, c}, {r, c - 1}, {r, c + 1}};
                for (auto& q : nb) {
                    if (q[0] < 0 || q[0] >= nr || q[1] < 0 || q[1] >= nc) continue;
                    const int qi = q[1] * nr + q[0];
                    if (mask[qi] && !keep[qi]) { keep[qi] = 1; stack.push_back(qi); }
                }
            }
        } else {
            for (size_t i = 0; i < img.size(); ++i) keep[i] = mask[i] ? 0 : 1;
        }
        for (int64_t e = e0; e < e1; ++e) {
            const int r = (int)(ir[e] % d1), c = (int)(ir[e] / d1);
            const bool inc = (r >= R0 && r <= R1 && c >= C0 && c <= C1);
            // outside the crop the mask is false: kept only in the "unlabelled remainder" case (those entries are zeros anyway)
            const bool kp = inc ? (keep[(size_t)(c - C0) * nr + (r - R0)] != 0) : !seeded;
            if (!kp) pr[e] = 0.0;
        }
    }
    return 0;
}

// circular_constraints(img) (endoscope/circular_constraints.m:1-54): on the bounding box of the non-zeros (skipped when it is
// a single row or column) zero the pixels whose gradient points away from the peak and that are below a third of it, keep
// the 4-connected component of the peak dilated by a 3 x 3 square, then a 3 x 3 median filter (zero padded, medfilt2).
// MATLAB gradient(): central differences inside, one-sided at the borders, spacing 1.
extern "C" int cnmfe_circular_constraints(int d1, int d2, int K, const int64_t* jc, const int64_t* ir, const double* pr,
                                          int64_t* out_jc, int64_t* out_ir, double* out_pr, int64_t cap) {
    if (d1 <= 0 || d2 <= 0 || K < 0 || !jc || !out_jc || (K > 0 && jc[K] > 0 && (!ir || !pr))) { set_error("cnmfe_circular_constraints: bad arguments"); return -1; }
    std::vector<double> img, tmp;
    std::vector<unsigned char> comp, dil;
    std::vector<int> stack;
    int64_t n = 0;
    out_jc[0] = 0;
    auto emit = [&](int64_t lin, double v) -> bool {
        if (v == 0.0) return true;
        if (n >= cap || !out_ir || !out_pr) return false;
        out_ir[n] = lin; out_pr[n] = v; ++n;
        return true;
    };
    for (int k = 0; k < K; ++k) {
        const int64_t e0 = jc[k], e1 = jc[k + 1];
        int r0 = d1, r1 = -1, c0 = d2, c1 = -1;
        for (int64_t e = e0; e < e1; ++e) {
            if (pr[e] == 0.0) continue;
            const int r = (int)(ir[e] % d1), c = (int)(ir[e] / d1);
            r0 = std::min(r0, r); r1 = std::max(r1, r); c0 = std::min(c0, c); c1 = std::max(c1, c);
        }
        const bool degenerate = (r1 < 0) || (r1 - r0 < 1) || (c1 - c0 < 1);       // empty, or a single row / column: unchanged
        if (degenerate) {
            for (int64_t e = e0; e < e1; ++e) if (!emit(ir[e], pr[e])) { set_error("cnmfe_circular_constraints: output too small"); return -1; }
            out_jc[k + 1] = n;
            continue;
        }
        const int nr = r1 - r0 + 1, nc = c1 - c0 + 1;
        img.assign((size_t)nr * nc, 0.0);
        for (int64_t e = e0; e < e1; ++e) {
            const int r = (int)(ir[e] % d1), c = (int)(ir[e] / d1);
            if (pr[e] != 0.0) img[(size_t)(c - c0) * nr + (r - r0)] = pr[e];
        }
        auto I = [&](int r, int c) -> double { return img[(size_t)c * nr + r]; };
        // peak: first maximum in column-major order
        double vmax = I(0, 0);
        int y0 = 0, x0 = 0;
        for (int c = 0; c < nc; ++c) for (int r = 0; r < nr; ++r) if (I(r, c) > vmax) { vmax = I(r, c); y0 = r; x0 = c; }
        tmp = img;
        for (int c = 0; c < nc; ++c)
            for (int r = 0; r < nr; ++r) {
                const double fx = (c == 0) ? I(r, 1) - I(r, 0) : (c == nc - 1) ? I(r, nc - 1) - I(r, nc - 2) : (I(r, c + 1) - I(r, c - 1)) / 2.0;
                const double fy = (r == 0) ? I(1, c) - I(0, c) : (r == nr - 1) ? I(nr - 1, c) - I(nr - 2, c) : (I(r + 1, c) - I(r - 1, c)) / 2.0;
                const double dot = fx * (double)(x0 - c) + fy * (double)(y0 - r);
                if (dot < 0.0 && I(r, c) < vmax / 3.0) tmp[(size_t)c * nr + r] = 0.0;
            }
        img.swap(tmp);
        // 4-connected component (non-zero pixels) of the peak, dilated by a 3 x 3 square
        comp.assign(img.size(), 0);
        if (I(y0, x0) != 0.0) {
            stack.clear(); stack.push_back(x0 * nr + y0); comp[stack.back()] = 1;
            while (!stack.empty()) {
                const int p = stack.back(); stack.pop_back();
                const int r = p % nr, c = p / nr;
                const int nb[4][2] = {{r - 1, c}, {r + 1, c}, {r, c - 1}, {r, c + 1}};
                for (auto& q : nb) {
                    if (q[0] < 0 || q[0] >= nr || q[1] < 0 || q[1] >= nc) continue;
                    const int qi = q[1] * nr + q[0];
                    if (img[qi] != 0.0 && !comp[qi]) { comp[qi] = 1; stack.push_back(qi); }
                }
            }
        } else {
            // l(ind_max) = 0: "l == 0" is the zero background (cannot happen for a positive peak; kept for completeness)
            for (size_t i = 0; i < img.size(); ++i) comp[i] = img[i] == 0.0 ? 1 : 0;
        }
        dil.assign(img.size(), 0);
        for (int c = 0; c < nc; ++c)
            for (int r = 0; r < nr; ++r) {
                bool any = false;
                for (int dc = -1; dc <= 1 && !any; ++dc)
                    for (int dr = -1; dr <= 1 && !any; ++dr) {
                        const int rr = r + dr, cc = c + dc;
                        if (rr >= 0 && rr < nr && cc >= 0 && cc < nc && comp[(size_t)cc * nr + rr]) any = true;
                    }
                dil[(size_t)c * nr + r] = any ? 1 : 0;
            }
        for (size_t i = 0; i < img.size(); ++i) if (!dil[i]) img[i] = 0.0;
        // medfilt2: 3 x 3 median, zeros outside the crop
        tmp.assign(img.size(), 0.0);
        for (int c = 0; c < nc; ++c)
            for (int r = 0; r < nr; ++r) {
                double w[9];
                int m = 0;
                for (int dc = -1; dc <= 1; ++dc)
                    for (int dr = -1; dr <= 1; ++dr) {
                        const int rr = r + dr, cc = c + dc;
                        w[m++] = (rr >= 0 && rr < nr && cc >= 0 && cc < nc) ? I(rr, cc) : 0.0;
                    }
                std::nth_element(w, w + 4, w + 9);
                tmp[(size_t)c * nr + r] = w[4];
            }
        for (int c = 0; c < nc; ++c)
            for (int r = 0; r < nr; ++r)
                if (!emit((int64_t)(c + c0) * d1 + (r + r0), tmp[(size_t)c * nr + r])) { set_error("cnmfe_circular_constraints: output too small"); return -1; }
        out_jc[k + 1] = n;
    }
    return 0;
}

// determine_search_location(A, 'ellipse', params): ellipse around the centre of mass, axes = principal components of the
// footprint with variances clamped to [min_size^2, max_size^2], expanded by `dist`.  Output: CSC pattern (sorted rows).
// out_ir must hold K * (2*ceil(dist*max_size) + 1)^2 entries (the ellipse never leaves that box).
extern "C" int cnmfe_search_location_ellipse(int d1, int d2, int K, const int64_t* jc, const int64_t* ir, const double* pr,
                                             double min_size, double max_size, double dist, int64_t* out_jc,
                                             int64_t* out_ir, int64_t cap) {
    if (d1 <= 0 || d2 <= 0 || K < 0 || !jc || !out_jc || (K > 0 && !out_ir)) { set_error("cnmfe_search_location_ellipse: bad arguments"); return -1; }
    if (!(dist > 0.0) || !std::isfinite(dist)) { set_error("cnmfe_search_location_ellipse: dist must be finite and positive (dist = Inf means the whole FOV: pass a full mask instead)"); return -1; }
    const int reach = (int)std::ceil(dist * std::max(max_size, min_size));
    int64_t n = 0;
    out_jc[0] = 0;
    for (int k = 0; k < K; ++k) {
        const int64_t e0 = jc[k], e1 = jc[k + 1];
        double s = 0.0, sx = 0.0, sy = 0.0;
        for (int64_t e = e0; e < e1; ++e) {
            const double a = pr ? pr[e] : 1.0;
            s += a;
            sx += a * (double)(ir[e] % d1 + 1);
            sy += a * (double)(ir[e] / d1 + 1);
        }
        double cx, cy, vxx = 0.0, vxy = 0.0, vyy = 0.0;
        if (s == 0.0) {
            // ind_empty: A(1, k) = 1  (determine_search_location.m:51-55)
            cx = 1.0; cy = 1.0;
        } else {
            cx = sx / s; cy = sy / s;
            // com.m:27-29 clamps
            if (cx < 0.0) cx = 0.0;
            if (cy < 0.0) cy = 0.0;
            if (cx > d1) cx = d1;
            if (cy > d2) cy = d2;
            for (int64_t e = e0; e < e1; ++e) {
                const double a = pr ? pr[e] : 1.0;
                const double dx = (double)(ir[e] % d1 + 1) - cx, dy = (double)(ir[e] / d1 + 1) - cy;
                vxx += a * dx * dx; vxy += a * dx * dy; vyy += a * dy * dy;
            }
            vxx /= s; vxy /= s; vyy /= s;
        }
        // eigen-decomposition of [[vxx, vxy], [vxy, vyy]], ascending eigenvalues (MATLAB eig of a symmetric matrix)
        const double half = 0.5 * (vxx + vyy), dif = 0.5 * (vxx - vyy);
        const double rad = std::sqrt(dif * dif + vxy * vxy);
        const double l1 = half - rad, l2 = half + rad;
        double v1x, v1y, v2x, v2y;
        if (vxy == 0.0) {
            if (vxx <= vyy) { v1x = 1; v1y = 0; v2x = 0; v2y = 1; } else { v1x = 0; v1y = 1; v2x = 1; v2y = 0; }
        } else {
            // eigenvector of l2: (vxy, l2 - vxx); of l1: orthogonal
            double ex = vxy, ey = l2 - vxx;
            const double nn = std::sqrt(ex * ex + ey * ey);
            ex /= nn; ey /= nn;
            v2x = ex; v2y = ey; v1x = -ey; v1y = ex;
        }
        const double d11 = std::min(max_size * max_size, std::max(min_size * min_size, l1));
        const double d22 = std::min(max_size * max_size, std::max(min_size * min_size, l2));
        const int rc = (int)std::floor(cx), cc = (int)std::floor(cy);     // 1-based centre, floor
        for (int c = std::max(1, cc - reach); c <= std::min(d2, cc + reach + 1); ++c)
            for (int r = std::max(1, rc - reach); r <= std::min(d1, rc + reach + 1); ++r) {
                const double dx = (double)r - cx, dy = (double)c - cy;
                const double p1 = dx * v1x + dy * v1y, p2 = dx * v2x + dy * v2y;
                if (std::sqrt(p1 * p1 / d11 + p2 * p2 / d22) <= dist) {
                    if (n >= cap) { set_error("cnmfe_search_location_ellipse: out_ir too small (%lld entries)", (long long)cap); return -1; }
                    out_ir[n++] = (int64_t)(c - 1) * d1 + (r - 1);
                }
            }
        out_jc[k + 1] = n;
    }
    return 0;
}

// determine_search_location(A, 'dilate', params) (utilities/determine_search_location.m:93-99): threshold_components
// (utilities/threshold_components.m: 3 x 3 median, keep the pixels carrying `nrgthr` of the energy, 3 x 3 closing, largest-
// energy 8-connected component; the LAST nb columns are copied unthresholded, as the reference does for its background
// columns) followed by a dilation with strel('disk', bSiz, 0).  Output: CSC pattern of IND.  cap: sum over neurons of
// (bbox height + 2 + 2 bSiz) * (bbox width + 2 + 2 bSiz) is always enough.
extern "C" int cnmfe_search_location_dilate(int d1, int d2, int K, const int64_t* jc, const int64_t* ir, const double* pr,
                                            double nrgthr, int nb, int bSiz, int64_t* out_jc, int64_t* out_ir, int64_t cap) {
    if (d1 <= 0 || d2 <= 0 || K < 0 || !jc || !out_jc || nb < 0 || bSiz < 0 || (K > 0 && jc[K] > 0 && (!ir || !pr))) { set_error("cnmfe_search_location_dilate: bad arguments"); return -1; }
    std::vector<double> img, med;
    std::vector<unsigned char> bw, tmpb, sel;
    std::vector<int> order, lab, stack;
    int64_t n = 0;
    out_jc[0] = 0;
    for (int k = 0; k < K; ++k) {
        const int64_t e0 = jc[k], e1 = jc[k + 1];
        int r0 = d1, r1 = -1, c0 = d2, c1 = -1;
        double colsum = 0.0;
        for (int64_t e = e0; e < e1; ++e) {
            colsum += pr[e];
            if (pr[e] == 0.0) continue;
            const int r = (int)(ir[e] % d1), c = (int)(ir[e] / d1);
            r0 = std::min(r0, r); r1 = std::max(r1, r); c0 = std::min(c0, c); c1 = std::max(c1, c);
        }
        const bool empty_col = (colsum == 0.0);                    // ind_empty: A(1, k) = 1  (:51-55)
        if (empty_col) { r0 = r1 = 0; c0 = c1 = 0; }
        if (r1 < 0) { out_jc[k + 1] = n; continue; }              // stored values cancel to a non-zero sum but no non-zero entry: nothing
        // crop with a margin of 2 (median spreads by 1, closing by 1 more), clipped to the FOV
        const int R0 = std::max(0, r0 - 2), R1 = std::min(d1 - 1, r1 + 2), C0 = std::max(0, c0 - 2), C1 = std::min(d2 - 1, c1 + 2);
        const int nr = R1 - R0 + 1, nc = C1 - C0 + 1;
        img.assign((size_t)nr * nc, 0.0);
        if (empty_col) img[(size_t)(0 - C0) * nr + (0 - R0)] = 1.0;
        for (int64_t e = e0; e < e1; ++e) {
            const int r = (int)(ir[e] % d1), c = (int)(ir[e] / d1);
            if (r >= R0 && r <= R1 && c >= C0 && c <= C1) img[(size_t)(c - C0) * nr + (r - R0)] += pr[e];   // (+= : pixel 1 of an "empty" column)
        }
        auto inside = [&](int r, int c) { return r >= 0 && r < nr && c >= 0 && c < nc; };
        const bool thresholded = (k < K - nb);
        if (thresholded) {
            // (i) medfilt2 3 x 3, zeros outside the image (pixels of the FOV outside the crop are zeros too)
            med.assign(img.size(), 0.0);
            for (int c = 0; c < nc; ++c)
                for (int r = 0; r < nr; ++r) {
                    double w[9]; int m = 0;
                    for (int dc = -1; dc <= 1; ++dc) for (int dr = -1; dr <= 1; ++dr) w[m++] = inside(r + dr, c + dc) ? img[(size_t)(c + dc) * nr + (r + dr)] : 0.0;
                    std::nth_element(w, w + 4, w + 9);
                    med[(size_t)c * nr + r] = w[4];
                }
            // (ii) energy threshold: ascending stable sort of the squares (zeros first, they add nothing), running sum
            order.clear();
            for (int i = 0; i < (int)med.size(); ++i) if (med[i] != 0.0) order.push_back(i);
            // FOV linear index order = (column, row) order; the crop index i = c * nr + r is monotone in it
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return med[a] * med[a] < med[b] * med[b]; });
            double total = 0.0;
            for (int i : order) total += med[i] * med[i];
            bw.assign(img.size(), 0);
            {
                const double level = (1.0 - nrgthr) * total;
                double run = 0.0;
                bool on = false;
                for (int i : order) {
                    run += med[i] * med[i];
                    if (!on && run > level) on = true;
                    if (on) bw[i] = 1;
                }
            }
            // (iii) imclose with a 3 x 3 square: dilation (outside = 0) then erosion (outside the FOV = 1; FOV pixels outside
            //       the crop are 0 after the dilation because the margin is 2)
            tmpb.assign(img.size(), 0);
            for (int c = 0; c < nc; ++c)
                for (int r = 0; r < nr; ++r) {
                    unsigned char v = 0;
                    for (int dc = -1; dc <= 1 && !v; ++dc) for (int dr = -1; dr <= 1 && !v; ++dr) if (inside(r + dr, c + dc) && bw[(size_t)(c + dc) * nr + (r + dr)]) v = 1;
                    tmpb[(size_t)c * nr + r] = v;
                }
            for (int c = 0; c < nc; ++c)
                for (int r = 0; r < nr; ++r) {
                    unsigned char v = 1;
                    for (int dc = -1; dc <= 1 && v; ++dc)
                        for (int dr = -1; dr <= 1 && v; ++dr) {
                            const int rr = r + dr, cc = c + dc;
                            const int fr = rr + R0, fc = cc + C0;
                            if (fr < 0 || fr >= d1 || fc < 0 || fc >= d2) continue;          // outside the FOV: ignored
                            if (!inside(rr, cc) || !tmpb[(size_t)cc * nr + rr]) v = 0;        // FOV pixel outside the crop: 0
                        }
                    bw[(size_t)c * nr + r] = v;
                }
            // (iv) 8-connected components in column-major discovery order; keep the one with the largest energy (first on ties)
            lab.assign(img.size(), 0);
            int ncomp = 0, best = 0;
            double best_e = -1.0;
            for (int i = 0; i < (int)bw.size(); ++i) {
                if (!bw[i] || lab[i]) continue;
                ++ncomp;
                stack.clear(); stack.push_back(i); lab[i] = ncomp;
                double e = 0.0;
                while (!stack.empty()) {
                    const int p = stack.back(); stack.pop_back();
                    e += med[p] * med[p];
                    const int r = p % nr, c = p / nr;
                    for (int dc = -1; dc <= 1; ++dc)
                        for (int dr = -1; dr <= 1; ++dr) {
                            if (!dr && !dc) continue;
                            if (!inside(r + dr, c + dc)) continue;
                            const int q = (c + dc) * nr + (r + dr);
                            if (bw[q] && !lab[q]) { lab[q] = ncomp; stack.push_back(q); }
                        }
                }
                if (e > best_e) { best_e = e; best = ncomp; }
            }
            sel.assign(img.size(), 0);
            for (size_t i = 0; i < img.size(); ++i) sel[i] = (best > 0 && lab[i] == best && med[i] > 0.0) ? 1 : 0;   // Ath > 0 after the dilation test
        } else {
            sel.assign(img.size(), 0);
            for (size_t i = 0; i < img.size(); ++i) sel[i] = img[i] > 0.0 ? 1 : 0;
        }
        // IND = imdilate(Ath, strel('disk', bSiz, 0)) > 0: a pixel is in if some selected positive pixel lies within the disk
        const int b = bSiz;
        const int DR0 = std::max(0, R0 - b), DR1 = std::min(d1 - 1, R1 + b), DC0 = std::max(0, C0 - b), DC1 = std::min(d2 - 1, C1 + b);
        for (int fc = DC0; fc <= DC1; ++fc)
            for (int fr = DR0; fr <= DR1; ++fr) {
                bool hit = false;
                for (int dc = -b; dc <= b && !hit; ++dc)
                    for (int dr = -b; dr <= b && !hit; ++dr) {
                        if (dr * dr + dc * dc > b * b) continue;
                        const int rr = fr + dr - R0, cc = fc + dc - C0;
                        if (inside(rr, cc) && sel[(size_t)cc * nr + rr]) hit = true;
                    }
                if (hit) {
                    if (n >= cap || !out_ir) { set_error("cnmfe_search_location_dilate: out_ir too small (%lld entries)", (long long)cap); return -1; }
                    out_ir[n++] = (int64_t)fc * d1 + fr;
                }
            }
        out_jc[k + 1] = n;
    }
    return 0;
}

// [l, c] = graph_connected_comp(sA)  (ca_source_extraction/utilities/graph_connected_comp.m:26 -> the reference's only native
// file, utilities/graph_conn_comp_mex.cpp): component labels 1..c of the nodes of a sparse adjacency matrix, numbered in the
// order of their smallest node.  A node's neighbours are the rows stored in its COLUMN (the reference follows the CSC lists
// only, so a non-symmetric matrix gives reachability from the seed and -- as there -- it is an error if that reaches a node
// labelled earlier).  Used by the merge routines (merge_components.m:47, MergeNeighbors.m:60, quickMerge.m:71, ...);
// SURVEY.md 8f row 4.
extern "C" int cnmfe_graph_conn_comp(int n, const int64_t* jc, const int64_t* ir, uint32_t* labels, int* ncomp) {
    if (n < 0 || !jc || !labels || !ncomp || (n > 0 && jc[n] > 0 && !ir)) { set_error("cnmfe_graph_conn_comp: bad arguments"); return -1; }
    std::fill(labels, labels + n, 0u);
    std::vector<int64_t> frontier;
    frontier.reserve((size_t)n);
    uint32_t cur = 0;
    for (int64_t seed = 0; seed < n; ++seed) {
        if (labels[seed]) continue;
        ++cur;
        labels[seed] = cur;
        frontier.clear();
        frontier.push_back(seed);
        for (size_t head = 0; head < frontier.size(); ++head) {
            const int64_t u = frontier[head];
            for (int64_t e = jc[u]; e < jc[u + 1]; ++e) {
                const int64_t v = ir[e];
                if (v < 0 || v >= n) { set_error("cnmfe_graph_conn_comp: row index %lld outside [0, %d)", (long long)v, n); return -1; }
                if (!labels[v]) { labels[v] = cur; frontier.push_back(v); }
                else if (labels[v] != cur) { set_error("cnmfe_graph_conn_comp: mixed labeling %u <-> %u (adjacency not symmetric)", labels[v], cur); return -1; }
            }
        }
    }
    *ncomp = (int)cur;
    return 0;
}
