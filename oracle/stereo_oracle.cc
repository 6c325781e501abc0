// stereo_oracle — TEST INFRASTRUCTURE: CPU restatement ("port") of Frame::ComputeStereoMatches
// (/root/reference/src/Frame.cc:501-675) over flat arrays: key points and descriptors of the left and right image, the
// two image pyramids (level interiors, as left in ORBextractor::mvImagePyramid), the scale tables, mb and mbf.
// Frame.cc cannot be compiled here (OpenCV, DBoW2, MapPoint ...) and the reference has no tests, so this port is pinned by
// an independent Python restatement (tests/golden/gen_stereo_golden.py -> committed fixtures): "unpinned by the reference".
// Build with -ffp-contract=off.
//
// Two places where the reference has undefined behaviour are given a definition (both unreachable for sane input):
//   * no accepted match at all: the reference reads vDistIdx[0] of an empty vector (:660); here nothing is removed;
//   * a right key point whose row band leaves the image (:516-522 would index vRowIndices out of range): rows are clipped.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

#include "../include/orbgpu.h"

namespace {
int descriptor_distance(const uint8_t* a, const uint8_t* b) {   // ORBmatcher.cc:1901-1917
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
}  // namespace

extern "C" void orbs_stereo_matches(const orbgpu_keypoint* keysL, const uint8_t* descL, int N, const orbgpu_keypoint* keysR,
                                    const uint8_t* descR, int Nr, const uint8_t* const* pyrL, const uint8_t* const* pyrR,
                                    const int32_t* level_w, const int32_t* level_h, const float* scale, const float* inv_scale,
                                    int n_levels, float mb, float mbf, float* uRight, float* depth, int32_t* sad_out) {
    const int TH_HIGH = 100, TH_LOW = 50;
    for (int i = 0; i < N; ++i) { uRight[i] = -1.0f; depth[i] = -1.0f; if (sad_out) sad_out[i] = -1; }
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;                         // :506
    const int nRows = level_h[0];                                        // :508
    std::vector<std::vector<size_t> > vRowIndices(nRows);                // :511-523
    for (int iR = 0; iR < Nr; iR++) {
        const float kpY = keysR[iR].y;
        const float r = 2.0f * scale[keysR[iR].octave];
        const int maxr = (int)ceil(kpY + r);
        const int minr = (int)floor(kpY - r);
        for (int yi = std::max(minr, 0); yi <= std::min(maxr, nRows - 1); yi++) vRowIndices[yi].push_back(iR);
    }
    const float minZ = mb, minD = 0, maxD = mbf / minZ;                   // :525-528
    std::vector<std::pair<int, int> > vDistIdx;
    for (int iL = 0; iL < N; iL++) {                                      // :532
        const orbgpu_keypoint& kpL = keysL[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y, uL = kpL.x;
        const int row = (int)vL;
        if (row < 0 || row >= nRows) continue;
        const std::vector<size_t>& vCandidates = vRowIndices[row];
        if (vCandidates.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH;
        size_t bestIdxR = 0;
        const uint8_t* dL = descL + (size_t)iL * 32;
        for (size_t iC = 0; iC < vCandidates.size(); iC++) {
            const size_t iR = vCandidates[iC];
            const orbgpu_keypoint& kpR = keysR[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = descriptor_distance(dL, descR + iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {                                       // :575 sub-pixel match by correlation
            const float uR0 = keysR[bestIdxR].x;
            const float scaleFactor = inv_scale[kpL.octave];
            const float scaleduL = roundf(kpL.x * scaleFactor);
            const float scaledvL = roundf(kpL.y * scaleFactor);
            const float scaleduR0 = roundf(uR0 * scaleFactor);
            const int w = 5, L = 5;
            const int lw = level_w[kpL.octave];
            const uint8_t* IL = pyrL[kpL.octave];
            const uint8_t* IR = pyrR[kpL.octave];
            const int cy = (int)scaledvL, cxL = (int)scaleduL, cxR0 = (int)scaleduR0;
            int bestSad = INT_MAX, bestincR = 0;
            float vDists[2 * L + 1];
            const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= lw) continue;
            const float cL = (float)IL[(size_t)cy * lw + cxL];
            for (int incR = -L; incR <= +L; incR++) {
                const float cR = (float)IR[(size_t)cy * lw + cxR0 + incR];
                double acc = 0;                                            // cv::norm(NORM_L1) of 32F accumulates in double
                for (int dy = -w; dy <= w; ++dy)
                    for (int dx = -w; dx <= w; ++dx) {
                        const float a = (float)IL[(size_t)(cy + dy) * lw + cxL + dx] - cL;
                        const float b = (float)IR[(size_t)(cy + dy) * lw + cxR0 + incR + dx] - cR;
                        acc += std::fabs(a - b);
                    }
                const float dist = (float)acc;
                if (dist < bestSad) { bestSad = (int)dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (bestincR == -L || bestincR == L) continue;
            const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = scale[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
                depth[iL] = mbf / disparity;
                uRight[iL] = bestuR;
                vDistIdx.push_back(std::pair<int, int>(bestSad, iL));
                if (sad_out) sad_out[iL] = bestSad;
            }
        }
    }
    if (vDistIdx.empty()) return;
    std::sort(vDistIdx.begin(), vDistIdx.end());                           // :659-674
    const float median = vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {
        if (vDistIdx[i].first < thDist) break;
        uRight[vDistIdx[i].second] = -1;
        depth[vDistIdx[i].second] = -1;
    }
}
