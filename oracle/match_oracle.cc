// match_oracle — TEST INFRASTRUCTURE: CPU restatement ("port") of the Hamming path of the reference ORBmatcher,
// over the same flat views the C ABI takes (include/orbgpu.h is included for the struct layouts only).
//
// Follows /root/reference/src/ORBmatcher.cc function by function (lines cited), plus the two Frame members the
// projection search calls: Frame::AssignFeaturesToGrid / PosInGrid (Frame.cc:232-247, :412-422) and
// Frame::GetFeaturesInArea (Frame.cc:353-410).  ORBmatcher.cc cannot be compiled here (it needs OpenCV, DBoW2 and
// the Frame/KeyFrame/MapPoint classes), and the reference ships no tests or golden vectors for it, so this
// restatement is pinned only by (a) DescriptorDistance against an independent popcount, (b) hand-checkable
// micro-cases in tests/test_oracle_matcher.py and (c) an independent Python restatement of the same functions
// (tests/golden/gen_matcher_golden.py -> committed fixtures).  Parity of these functions is therefore "unpinned by
// the reference" (DESIGN.md §Oracle).  Build with -ffp-contract=off.
#include <climits>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/orbgpu.h"

namespace {

const int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;  // ORBmatcher.cc:37-39

// ORBmatcher::DescriptorDistance, ORBmatcher.cc:1901-1917 (SWAR bit count over eight 32-bit words)
int descriptor_distance(const uint8_t* a, const uint8_t* b) {
    int32_t pa[8], pb[8];
    memcpy(pa, a, 32);
    memcpy(pb, b, 32);
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        unsigned int v = pa[i] ^ pb[i];
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

// ORBmatcher::ComputeThreeMaxima, ORBmatcher.cc:1854-1895
void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

// the rotation-histogram idiom shared by the BoW and triangulation searches (e.g. :718-728)
int rot_bin(float a1, float a2) {
    const float factor = 1.0f / HISTO_LENGTH;
    float rot = a1 - a2;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)round(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

struct Grid {  // Frame::mGrid, Frame.h:161
    std::vector<int> cell[ORBGPU_GRID_COLS][ORBGPU_GRID_ROWS];
};

// Frame::AssignFeaturesToGrid + PosInGrid, Frame.cc:232-247, :412-422
void assign_grid(const orbgpu_keypoint* k, int n, const float* g, Grid& G) {
    for (int i = 0; i < n; i++) {
        const int px = (int)round((k[i].x - g[0]) * g[2]);
        const int py = (int)round((k[i].y - g[1]) * g[3]);
        if (px < 0 || px >= ORBGPU_GRID_COLS || py < 0 || py >= ORBGPU_GRID_ROWS) continue;
        G.cell[px][py].push_back(i);
    }
}

// Frame::GetFeaturesInArea, Frame.cc:353-410
void features_in_area(const Grid& G, const orbgpu_keypoint* k, const float* g, float x, float y, float r, int minLevel,
                      int maxLevel, std::vector<int>& out) {
    out.clear();
    const int nMinCellX = std::max(0, (int)floor((x - g[0] - r) * g[2]));
    if (nMinCellX >= ORBGPU_GRID_COLS) return;
    const int nMaxCellX = std::min((int)ORBGPU_GRID_COLS - 1, (int)ceil((x - g[0] + r) * g[2]));
    if (nMaxCellX < 0) return;
    const int nMinCellY = std::max(0, (int)floor((y - g[1] - r) * g[3]));
    if (nMinCellY >= ORBGPU_GRID_ROWS) return;
    const int nMaxCellY = std::min((int)ORBGPU_GRID_ROWS - 1, (int)ceil((y - g[1] + r) * g[3]));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const std::vector<int>& vCell = G.cell[ix][iy];
            for (size_t j = 0; j < vCell.size(); j++) {
                const orbgpu_keypoint& kp = k[vCell[j]];
                if (bCheckLevels) {
                    if (kp.octave < minLevel) continue;
                    if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                }
                const float distx = kp.x - x, disty = kp.y - y;
                if (fabs(distx) < r && fabs(disty) < r) out.push_back(vCell[j]);
            }
        }
}

// ORBmatcher::CheckDistEpipolarLine, ORBmatcher.cc:173-196 (F12 row major)
bool check_epipolar(const orbgpu_keypoint& kp1, const orbgpu_keypoint& kp2, const float* F, const float* sigma2) {
    const float a = kp1.x * F[0] + kp1.y * F[3] + F[6];
    const float b = kp1.x * F[1] + kp1.y * F[4] + F[7];
    const float c = kp1.x * F[2] + kp1.y * F[5] + F[8];
    const float num = a * kp2.x + b * kp2.y + c;
    const float den = a * a + b * b;
    if (den == 0) return false;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * sigma2[kp2.octave];
}

}  // namespace

extern "C" {

int orbm_hamming(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th), ORBmatcher.cc:59-155
void orbm_search_by_projection(const orbgpu_frame_set* fs, const orbgpu_mappoint_set* mp, const float* scale, int n_levels,
                               float th, float nnratio, int32_t* kp_match, int32_t* mp_best_idx, int32_t* mp_best_dist,
                               int32_t* mp_second_dist, int32_t* nmatches_out) {
    for (int f = 0; f < fs->n_frames; ++f) {
        const int k0 = fs->kp_off[f], n = fs->kp_off[f + 1] - k0;
        const orbgpu_keypoint* keys = fs->keys_un + k0;
        const float* g = fs->grid + 4 * f;
        Grid G;
        assign_grid(keys, n, g, G);
        std::vector<uint8_t> state(n, 0);
        if (fs->kp_flags) for (int i = 0; i < n; ++i) state[i] = fs->kp_flags[k0 + i];
        for (int i = 0; i < n; ++i) if (kp_match) kp_match[k0 + i] = -1;
        int nmatches = 0;
        const bool bFactor = th != 1.0;
        std::vector<int> vIndices;
        for (int q = mp->mp_off[f]; q < mp->mp_off[f + 1]; ++q) {
            if (mp_best_idx) mp_best_idx[q] = -1;
            if (mp_best_dist) mp_best_dist[q] = 256;
            if (mp_second_dist) mp_second_dist[q] = 256;
            if (!(mp->flags[q] & 1)) continue;   // mbTrackInView
            if (mp->flags[q] & 2) continue;      // isBad()
            const int lvl = mp->level[q];
            float r = mp->view_cos[q] > 0.998 ? 2.5f : 4.0f;  // RadiusByViewingCos, :157-163
            if (bFactor) r *= th;
            features_in_area(G, keys, g, mp->proj_x[q], mp->proj_y[q], r * scale[lvl], lvl - 1, lvl, vIndices);
            if (vIndices.empty()) continue;
            const uint8_t* d = mp->desc + (size_t)q * 32;
            int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
            for (size_t j = 0; j < vIndices.size(); ++j) {
                const int idx = vIndices[j];
                if (state[idx] == 1) continue;  // mvpMapPoints[idx] with Observations()>0, :108-110
                if (fs->u_right && fs->u_right[k0 + idx] > 0) {
                    const float er = fabs(mp->proj_xr[q] - fs->u_right[k0 + idx]);
                    if (er > r * scale[lvl]) continue;
                }
                const int dist = descriptor_distance(d, fs->desc + (size_t)(k0 + idx) * 32);
                if (dist < bestDist) {
                    bestDist2 = bestDist; bestDist = dist;
                    bestLevel2 = bestLevel; bestLevel = keys[idx].octave;
                    bestIdx = idx;
                } else if (dist < bestDist2) {
                    bestLevel2 = keys[idx].octave;
                    bestDist2 = dist;
                }
            }
            if (mp_best_idx) mp_best_idx[q] = bestIdx;
            if (mp_best_dist) mp_best_dist[q] = bestDist;
            if (mp_second_dist) mp_second_dist[q] = bestDist2;
            if (bestDist <= TH_HIGH) {
                if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
                state[bestIdx] = (mp->flags[q] & 4) ? 1 : 2;   // F.mvpMapPoints[bestIdx] = pMP
                if (kp_match) kp_match[k0 + bestIdx] = q - mp->mp_off[f];
                nmatches++;
            }
        }
        if (nmatches_out) nmatches_out[f] = nmatches;
    }
}

// The search loop of SearchByProjection(Frame&, const Frame&, th, bMono) (ORBmatcher.cc:1581-1684) and of
// SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist) (:1760-1832) over already projected queries.
void orbm_search_windowed(const orbgpu_frame_set* fs, const orbgpu_window_query_set* qs, int th_dist, int skip_any, int check_orientation,
                          int32_t* kp_match, int32_t* q_best_idx, int32_t* q_best_dist, int32_t* nmatches_out) {
    for (int f = 0; f < fs->n_frames; ++f) {
        const int k0 = fs->kp_off[f], n = fs->kp_off[f + 1] - k0;
        const orbgpu_keypoint* keys = fs->keys_un + k0;
        const float* g = fs->grid + 4 * f;
        Grid G;
        assign_grid(keys, n, g, G);
        std::vector<uint8_t> state(n, 0);   // 0: NULL, 1: MapPoint with observations, 2: MapPoint without
        if (fs->kp_flags) for (int i = 0; i < n; ++i) state[i] = fs->kp_flags[k0 + i];
        for (int i = 0; i < n; ++i) if (kp_match) kp_match[k0 + i] = -1;
        int nmatches = 0;
        std::vector<int> rotHist[HISTO_LENGTH];
        std::vector<int> vIndices2;
        for (int q = qs->q_off[f]; q < qs->q_off[f + 1]; ++q) {
            if (q_best_idx) q_best_idx[q] = -1;
            if (q_best_dist) q_best_dist[q] = 256;
            if (!(qs->flags[q] & 1)) continue;
            const float u = qs->u[q], v = qs->v[q], radius = qs->radius[q];
            features_in_area(G, keys, g, u, v, radius, qs->min_level[q], qs->max_level[q], vIndices2);
            if (vIndices2.empty()) continue;
            const uint8_t* dMP = qs->desc + (size_t)q * 32;
            int bestDist = 256, bestIdx2 = -1;
            for (size_t j = 0; j < vIndices2.size(); ++j) {
                const int i2 = vIndices2[j];
                if (state[i2] == 1 || (skip_any && state[i2] != 0)) continue;   // :1619-1621 / :1776-1777
                if (qs->ur && fs->u_right && fs->u_right[k0 + i2] > 0) {         // :1624-1630
                    const float er = fabs(qs->ur[q] - fs->u_right[k0 + i2]);
                    if (er > radius) continue;
                }
                const int dist = descriptor_distance(dMP, fs->desc + (size_t)(k0 + i2) * 32);
                if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
            }
            if (q_best_idx) q_best_idx[q] = bestIdx2;
            if (q_best_dist) q_best_dist[q] = bestDist;
            if (bestDist <= th_dist) {
                state[bestIdx2] = (qs->flags[q] & 4) ? 1 : 2;                   // CurrentFrame.mvpMapPoints[bestIdx2] = pMP
                if (kp_match) kp_match[k0 + bestIdx2] = q - qs->q_off[f];
                nmatches++;
                if (check_orientation) rotHist[rot_bin(qs->angle[q], keys[bestIdx2].angle)].push_back(bestIdx2);
            }
        }
        if (check_orientation) {
            int ind1 = -1, ind2 = -1, ind3 = -1;
            three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
            for (int i = 0; i < HISTO_LENGTH; i++) {
                if (i == ind1 || i == ind2 || i == ind3) continue;
                for (size_t j = 0; j < rotHist[i].size(); j++) {
                    if (kp_match) kp_match[k0 + rotHist[i][j]] = -2;            // reset to NULL
                    nmatches--;
                }
            }
        }
        if (nmatches_out) nmatches_out[f] = nmatches;
    }
}

// ORBmatcher::SearchForTriangulation, ORBmatcher.cc:783-975
void orbm_search_for_triangulation(const orbgpu_frame_set* s1, const orbgpu_frame_set* s2, int n_pairs, const int32_t* idx1v,
                                   const int32_t* idx2v, const float* f12, const float* epipole, const float* scale,
                                   const float* sigma2, int n_levels, int only_stereo, int check_orientation,
                                   const int64_t* match_off, int32_t* match12, int32_t* match_dist, int32_t* nmatches_out) {
    for (int p = 0; p < n_pairs; ++p) {
        const int fa = idx1v[p], fb = idx2v[p];
        const int ka = s1->kp_off[fa], na = s1->kp_off[fa + 1] - ka;
        const int kb = s2->kp_off[fb];
        const orbgpu_keypoint* K1 = s1->keys_un + ka;
        const orbgpu_keypoint* K2 = s2->keys_un + kb;
        const float ex = epipole[2 * p], ey = epipole[2 * p + 1];
        const float* F = f12 + 9 * p;
        int nmatches = 0;
        std::vector<int> vMatches12(na, -1), vDist(na, -1);
        std::vector<int> rotHist[HISTO_LENGTH];
        int a = s1->fv_node_off[fa], aend = s1->fv_node_off[fa + 1];
        int b = s2->fv_node_off[fb], bend = s2->fv_node_off[fb + 1];
        while (a != aend && b != bend) {
            if (s1->fv_node_id[a] == s2->fv_node_id[b]) {
                for (int i1 = s1->fv_feat_off[a]; i1 < s1->fv_feat_off[a + 1]; i1++) {
                    const int idx1 = s1->fv_feat[i1];
                    if (s1->kp_flags && (s1->kp_flags[ka + idx1] & 1)) continue;   // pMP1 set, :846
                    const bool bStereo1 = s1->u_right && s1->u_right[ka + idx1] >= 0;
                    if (only_stereo && !bStereo1) continue;
                    const orbgpu_keypoint& kp1 = K1[idx1];
                    const uint8_t* d1 = s1->desc + (size_t)(ka + idx1) * 32;
                    int bestDist = TH_LOW, bestIdx2 = -1;
                    for (int i2 = s2->fv_feat_off[b]; i2 < s2->fv_feat_off[b + 1]; i2++) {
                        const int idx2 = s2->fv_feat[i2];
                        if (s2->kp_flags && (s2->kp_flags[kb + idx2] & 1)) continue;  // vbMatched2 is never set, :868
                        const bool bStereo2 = s2->u_right && s2->u_right[kb + idx2] >= 0;
                        if (only_stereo && !bStereo2) continue;
                        const int dist = descriptor_distance(d1, s2->desc + (size_t)(kb + idx2) * 32);
                        if (dist > TH_LOW || dist > bestDist) continue;
                        const orbgpu_keypoint& kp2 = K2[idx2];
                        if (!bStereo1 && !bStereo2) {
                            const float distex = ex - kp2.x, distey = ey - kp2.y;
                            if (distex * distex + distey * distey < 100 * scale[kp2.octave]) continue;
                        }
                        if (check_epipolar(kp1, kp2, F, sigma2)) { bestIdx2 = idx2; bestDist = dist; }
                    }
                    if (bestIdx2 >= 0) {
                        vMatches12[idx1] = bestIdx2;
                        vDist[idx1] = bestDist;
                        nmatches++;
                        if (check_orientation) rotHist[rot_bin(kp1.angle, K2[bestIdx2].angle)].push_back(idx1);
                    }
                }
                a++; b++;
            } else if (s1->fv_node_id[a] < s2->fv_node_id[b]) {
                while (a != aend && s1->fv_node_id[a] < s2->fv_node_id[b]) a++;   // lower_bound, :932
            } else {
                while (b != bend && s2->fv_node_id[b] < s1->fv_node_id[a]) b++;
            }
        }
        if (check_orientation) {
            int ind1 = -1, ind2 = -1, ind3 = -1;
            three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
            for (int i = 0; i < HISTO_LENGTH; i++) {
                if (i == ind1 || i == ind2 || i == ind3) continue;
                for (size_t j = 0; j < rotHist[i].size(); j++) { vMatches12[rotHist[i][j]] = -1; vDist[rotHist[i][j]] = -1; nmatches--; }
            }
        }
        for (int i = 0; i < na; ++i) {
            match12[match_off[p] + i] = vMatches12[i];
            if (match_dist) match_dist[match_off[p] + i] = vDist[i];
        }
        if (nmatches_out) nmatches_out[p] = nmatches;
    }
}

// ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, ...) :635-768 and SearchByBoW(KeyFrame*, Frame&, ...) :211-344
void orbm_search_by_bow(const orbgpu_frame_set* s1, const orbgpu_frame_set* s2, int n_pairs, const int32_t* idx1v,
                        const int32_t* idx2v, float nnratio, int check_orientation, int th_low, int th_inclusive,
                        int require_mp2, const int64_t* match_off, int32_t* match12, int32_t* match_dist,
                        int32_t* nmatches_out) {
    for (int p = 0; p < n_pairs; ++p) {
        const int fa = idx1v[p], fb = idx2v[p];
        const int ka = s1->kp_off[fa], na = s1->kp_off[fa + 1] - ka;
        const int kb = s2->kp_off[fb], nb = s2->kp_off[fb + 1] - kb;
        const orbgpu_keypoint* K1 = s1->keys_un + ka;
        const orbgpu_keypoint* K2 = s2->keys_un + kb;
        std::vector<int> vMatches12(na, -1), vDist(na, -1);
        std::vector<bool> vbMatched2(nb, false);
        std::vector<int> rotHist[HISTO_LENGTH];
        int nmatches = 0;
        int a = s1->fv_node_off[fa], aend = s1->fv_node_off[fa + 1];
        int b = s2->fv_node_off[fb], bend = s2->fv_node_off[fb + 1];
        while (a != aend && b != bend) {
            if (s1->fv_node_id[a] == s2->fv_node_id[b]) {
                for (int i1 = s1->fv_feat_off[a]; i1 < s1->fv_feat_off[a + 1]; i1++) {
                    const int idx1 = s1->fv_feat[i1];
                    if (!(s1->kp_flags && (s1->kp_flags[ka + idx1] & 1))) continue;   // !pMP1 || isBad, :673-677
                    const uint8_t* d1 = s1->desc + (size_t)(ka + idx1) * 32;
                    int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                    for (int i2 = s2->fv_feat_off[b]; i2 < s2->fv_feat_off[b + 1]; i2++) {
                        const int idx2 = s2->fv_feat[i2];
                        if (vbMatched2[idx2]) continue;
                        if (require_mp2 && !(s2->kp_flags && (s2->kp_flags[kb + idx2] & 1))) continue;
                        const int dist = descriptor_distance(d1, s2->desc + (size_t)(kb + idx2) * 32);
                        if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
                        else if (dist < bestDist2) { bestDist2 = dist; }
                    }
                    const bool pass = th_inclusive ? (bestDist1 <= th_low) : (bestDist1 < th_low);
                    if (pass && static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
                        vMatches12[idx1] = bestIdx2;
                        vDist[idx1] = bestDist1;
                        vbMatched2[bestIdx2] = true;
                        if (check_orientation) rotHist[rot_bin(K1[idx1].angle, K2[bestIdx2].angle)].push_back(idx1);
                        nmatches++;
                    }
                }
                a++; b++;
            } else if (s1->fv_node_id[a] < s2->fv_node_id[b]) {
                while (a != aend && s1->fv_node_id[a] < s2->fv_node_id[b]) a++;
            } else {
                while (b != bend && s2->fv_node_id[b] < s1->fv_node_id[a]) b++;
            }
        }
        if (check_orientation) {
            int ind1 = -1, ind2 = -1, ind3 = -1;
            three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
            for (int i = 0; i < HISTO_LENGTH; i++) {
                if (i == ind1 || i == ind2 || i == ind3) continue;
                for (size_t j = 0; j < rotHist[i].size(); j++) { vMatches12[rotHist[i][j]] = -1; vDist[rotHist[i][j]] = -1; nmatches--; }
            }
        }
        for (int i = 0; i < na; ++i) {
            match12[match_off[p] + i] = vMatches12[i];
            if (match_dist) match_dist[match_off[p] + i] = vDist[i];
        }
        if (nmatches_out) nmatches_out[p] = nmatches;
    }
}

// CPU baseline driver: the pairs are split over nthreads std::threads (one pair at a time per thread).
}  // extern "C"

#include <chrono>
#include <thread>
extern "C" double orbm_bench_bow(const orbgpu_frame_set* s1, const orbgpu_frame_set* s2, int n_pairs, const int32_t* idx1v,
                                 const int32_t* idx2v, float nnratio, int check_orientation, int th_low, int th_inclusive,
                                 int require_mp2, const int64_t* match_off, int32_t* match12, int32_t* nmatches, int nthreads) {
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([=]() {
            for (int p = t; p < n_pairs; p += nthreads) {
                // a one-pair view: same sets, shifted pair arrays
                orbm_search_by_bow(s1, s2, 1, idx1v + p, idx2v + p, nnratio, check_orientation, th_low, th_inclusive, require_mp2,
                                   match_off + p, match12, nullptr, nmatches + p);
            }
        });
    for (auto& x : th) x.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// Frame::isInFrustum (Frame.cc:274-342) + MapPoint::PredictScale (MapPoint.cc:421-436) for every map point of every frame.
// cam[f] = Rcw[9], tcw[3], Ow[3], fx, fy, cx, cy, mbf, mnMinX, mnMaxX, mnMinY, mnMaxY (24 floats).  The cv::Mat expressions
// follow cv2 4.13 (oracle/cvlite.cc cvl_gemm3_f32 / cvl_norm3_f32 / cvl_dot3_f32, pinned by tests/golden/cvsmall_golden.npz).
extern "C" void cvl_gemm3_f32(const float*, const float*, const float*, float*);
extern "C" double cvl_norm3_f32(const float*);
extern "C" double cvl_dot3_f32(const float*, const float*);
extern "C" void orbm_is_in_frustum(int n_frames, const float* cam, float log_scale_factor, int n_levels, float viewing_cos_limit,
                                   const int32_t* mp_off, const float* world_pos, const float* normal, const float* min_dist_inv,
                                   const float* max_dist_inv, const float* max_distance, uint8_t* in_view, float* proj_x, float* proj_y,
                                   float* proj_xr, int32_t* level, float* view_cos) {
    for (int f = 0; f < n_frames; ++f) {
        const float* c = cam + 24 * f;
        const float fx = c[15], fy = c[16], cx = c[17], cy = c[18], mbf = c[19], minX = c[20], maxX = c[21], minY = c[22], maxY = c[23];
        for (int q = mp_off[f]; q < mp_off[f + 1]; ++q) {
            in_view[q] = 0;                                   // pMP->mbTrackInView = false (:277)
            proj_x[q] = proj_y[q] = proj_xr[q] = view_cos[q] = 0.f;
            level[q] = 0;
            const float* P = world_pos + 3 * q;
            float Pc[3];
            cvl_gemm3_f32(c, P, c + 9, Pc);                   // mRcw*P+mtcw (:285)
            const float PcX = Pc[0], PcY = Pc[1], PcZ = Pc[2];
            if (PcZ < 0.0f) continue;                         // :292
            const float invz = 1.0f / PcZ;
            const float u = fx * PcX * invz + cx;
            const float v = fy * PcY * invz + cy;
            if (u < minX || u > maxX) continue;               // :301-304
            if (v < minY || v > maxY) continue;
            const float PO[3] = {P[0] - c[12], P[1] - c[13], P[2] - c[14]};   // P-mOw (:312)
            const float dist = (float)cvl_norm3_f32(PO);
            if (dist < min_dist_inv[q] || dist > max_dist_inv[q]) continue;    // :315
            const float viewCos = (float)(cvl_dot3_f32(PO, normal + 3 * q) / dist);   // :322
            if (viewCos < viewing_cos_limit) continue;
            const float ratio = max_distance[q] / dist;       // MapPoint.cc:426
            int nScale = (int)std::ceil(std::log(ratio) / log_scale_factor);   // float overloads, as in the reference (:429)
            if (nScale < 0) nScale = 0;
            else if (nScale >= n_levels) nScale = n_levels - 1;
            in_view[q] = 1;
            proj_x[q] = u;
            proj_xr[q] = u - mbf * invz;
            proj_y[q] = v;
            level[q] = nScale;
            view_cos[q] = viewCos;
        }
    }
}

// MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:247-316) for a batch of map points: point p owns the observed
// descriptors [obs_off[p], obs_off[p+1]) (those of its non-bad key frames, in std::map<KeyFrame*, size_t> iteration order).
// best_idx = the row with the least median distance to the others (first wins), -1 for a point without descriptors.
extern "C" void orbm_distinctive_descriptors(int n_points, const int32_t* obs_off, const uint8_t* desc, int32_t* best_idx, int32_t* best_median) {
    for (int p = 0; p < n_points; ++p) {
        const int N = obs_off[p + 1] - obs_off[p];
        best_idx[p] = -1;
        best_median[p] = INT_MAX;
        if (N <= 0) continue;
        const uint8_t* D = desc + (size_t)obs_off[p] * 32;
        std::vector<std::vector<float>> Distances(N, std::vector<float>(N, 0.f));
        for (int i = 0; i < N; ++i) {
            Distances[i][i] = 0;
            for (int j = i + 1; j < N; ++j) {
                const int d = descriptor_distance(D + (size_t)i * 32, D + (size_t)j * 32);
                Distances[i][j] = d;
                Distances[j][i] = d;
            }
        }
        int BestMedian = INT_MAX, BestIdx = 0;
        for (int i = 0; i < N; ++i) {
            std::vector<int> vDists(Distances[i].begin(), Distances[i].end());
            std::sort(vDists.begin(), vDists.end());
            const int median = vDists[0.5 * (N - 1)];
            if (median < BestMedian) { BestMedian = median; BestIdx = i; }
        }
        best_idx[p] = BestIdx;
        best_median[p] = BestMedian;
    }
}

// The candidate loops of Fuse (ORBmatcher.cc:1051-1112, :1211-1246) and SearchBySim3 (:1363-1401, :1443-1481): independent
// queries, best candidate only; optional chi-square gate of Fuse(KF, vpMapPoints, th) (:1077-1102).
extern "C" void orbm_search_window_best(const orbgpu_frame_set* fs, const orbgpu_window_query_set* qs, const float* inv_sigma2, int skip_flagged,
                                        int32_t* q_best_idx, int32_t* q_best_dist) {
    for (int f = 0; f < fs->n_frames; ++f) {
        const int k0 = fs->kp_off[f], n = fs->kp_off[f + 1] - k0;
        const orbgpu_keypoint* keys = fs->keys_un + k0;
        const float* g = fs->grid + 4 * f;
        Grid G;
        assign_grid(keys, n, g, G);
        std::vector<int> vIndices;
        for (int q = qs->q_off[f]; q < qs->q_off[f + 1]; ++q) {
            q_best_idx[q] = -1;
            q_best_dist[q] = 256;
            if (!(qs->flags[q] & 1)) continue;
            const float u = qs->u[q], v = qs->v[q];
            features_in_area(G, keys, g, u, v, qs->radius[q], qs->min_level[q], qs->max_level[q], vIndices);
            int bestDist = 256, bestIdx = -1;
            for (size_t j = 0; j < vIndices.size(); ++j) {
                const int idx = vIndices[j];
                if (skip_flagged && fs->kp_flags && fs->kp_flags[k0 + idx] != 0) continue;
                if (inv_sigma2) {
                    const orbgpu_keypoint& kp = keys[idx];
                    const float kur = fs->u_right ? fs->u_right[k0 + idx] : -1.f;
                    const float ex = u - kp.x, ey = v - kp.y;
                    if (kur >= 0) {
                        const float er = qs->ur[q] - kur;
                        const float e2 = ex * ex + ey * ey + er * er;
                        if (e2 * inv_sigma2[kp.octave] > 7.8) continue;
                    } else {
                        const float e2 = ex * ex + ey * ey;
                        if (e2 * inv_sigma2[kp.octave] > 5.99) continue;
                    }
                }
                const int dist = descriptor_distance(qs->desc + (size_t)q * 32, fs->desc + (size_t)(k0 + idx) * 32);
                if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
            }
            q_best_idx[q] = bestIdx;
            q_best_dist[q] = bestDist;
        }
    }
}

// ORBmatcher::SearchForInitialization, ORBmatcher.cc:493-632 (queries = key points of F1, frames = F2)
extern "C" void orbm_search_for_initialization(const orbgpu_frame_set* fs, const orbgpu_window_query_set* qs, float nnratio, int check_orientation,
                                               int32_t* match12, int32_t* nmatches_out) {
    for (int f = 0; f < fs->n_frames; ++f) {
        const int k0 = fs->kp_off[f], n = fs->kp_off[f + 1] - k0;
        const orbgpu_keypoint* keys = fs->keys_un + k0;
        const float* g = fs->grid + 4 * f;
        Grid G;
        assign_grid(keys, n, g, G);
        const int q0 = qs->q_off[f], nq = qs->q_off[f + 1] - q0;
        int nmatches = 0;
        std::vector<int> vnMatches12(nq, -1), vMatchedDistance(n, INT_MAX), vnMatches21(n, -1);
        std::vector<int> rotHist[HISTO_LENGTH];
        std::vector<int> vIndices2;
        for (int i1 = 0; i1 < nq; ++i1) {
            const int q = q0 + i1;
            if (!(qs->flags[q] & 1)) continue;   // level1 > 0
            features_in_area(G, keys, g, qs->u[q], qs->v[q], qs->radius[q], qs->min_level[q], qs->max_level[q], vIndices2);
            if (vIndices2.empty()) continue;
            const uint8_t* d1 = qs->desc + (size_t)q * 32;
            int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
            for (size_t j = 0; j < vIndices2.size(); ++j) {
                const int i2 = vIndices2[j];
                const int dist = descriptor_distance(d1, fs->desc + (size_t)(k0 + i2) * 32);
                if (vMatchedDistance[i2] <= dist) continue;
                if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
                else if (dist < bestDist2) bestDist2 = dist;
            }
            if (bestDist <= TH_LOW) {
                if (bestDist < (float)bestDist2 * nnratio) {
                    if (vnMatches21[bestIdx2] >= 0) { vnMatches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                    vnMatches12[i1] = bestIdx2;
                    vnMatches21[bestIdx2] = i1;
                    vMatchedDistance[bestIdx2] = bestDist;
                    nmatches++;
                    if (check_orientation) rotHist[rot_bin(qs->angle[q], keys[bestIdx2].angle)].push_back(i1);
                }
            }
        }
        if (check_orientation) {
            int ind1 = -1, ind2 = -1, ind3 = -1;
            three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
            for (int i = 0; i < HISTO_LENGTH; i++) {
                if (i == ind1 || i == ind2 || i == ind3) continue;
                for (size_t j = 0; j < rotHist[i].size(); j++) {
                    const int idx1 = rotHist[i][j];
                    if (vnMatches12[idx1] >= 0) { vnMatches12[idx1] = -1; nmatches--; }
                }
            }
        }
        for (int i1 = 0; i1 < nq; ++i1) match12[q0 + i1] = vnMatches12[i1];
        if (nmatches_out) nmatches_out[f] = nmatches;
    }
}
