// orb_oracle — TEST INFRASTRUCTURE: CPU restatement ("port") of the reference ORB extractor.
//
// Restates, over flat arrays and without any cv:: type, what /root/reference/src/ORBextractor.cc
// computes; every function cites the lines it follows.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline leg may load this library; the product (liborbgpu.so) never does.
//
// Pinning: (1) the OpenCV primitives it calls (cvlite) are checked against cv2 4.13.0 golden
// vectors; (2) the glue restated here is checked for equality against the verbatim compile of the
// reference source (oracle/_ref/liborbref.so, see oracle/Makefile) and against committed fixtures
// produced by genuine cv2 primitives + an independent Python glue (tests/golden/).  The reference
// itself ships no tests or golden vectors, so parity is pinned only by those.
//
// One documented deviation from the reference: the pointer-valued tie-break of the node sort at
// ORBextractor.cc:684 is replaced by "ties keep creation order" (SURVEY.md §7.3 #1).  Build flags:
// -ffp-contract=off (the rounded rotations at :118-120 must not be fused).
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <list>
#include <vector>
#include <algorithm>

#include "cvlite.h"
#include "../include/orbgpu_pattern.inc"

namespace {

const int PATCH_SIZE = 31;        // ORBextractor.cc:72
const int HALF_PATCH_SIZE = 15;   // :73
const int EDGE_THRESHOLD = 19;    // :74

const int8_t kPatX[512] = {ORB_PATTERN_X_INIT};
const int8_t kPatY[512] = {ORB_PATTERN_Y_INIT};

struct Kp {           // byte-compatible with cv::KeyPoint (28 B)
    float x, y, size, angle, response;
    int32_t octave, class_id;
};

struct Level {
    int w = 0, h = 0, stride = 0;            // interior size; stride of the bordered buffer
    std::vector<uint8_t> buf;                // (w+38) x (h+38), interior at (19,19)
    std::vector<uint8_t> blurred;            // w x h
    std::vector<Kp> candidates;              // FAST output in emission order, coords relative to minBorder
    std::vector<Kp> keypoints;               // after octree + orientation, level coordinates
    uint8_t* interior() { return buf.data() + (size_t)EDGE_THRESHOLD * stride + EDGE_THRESHOLD; }
};

inline int cv_round_f(float v) { return (int)lrintf(v); }
inline int cv_round_d(double v) { return (int)lrint(v); }

struct Node {                      // ORBextractor.h:32-43
    std::vector<int> keys;         // indices into the candidate array, emission order preserved
    int ulx, uly, urx, ury, blx, bly, brx, bry;
    bool noMore = false;
    std::list<Node>::iterator self;
    long created = 0;              // creation counter: the oracle's tie-break (replaces the pointer)
};

struct Oracle {
    int nfeatures, nlevels, iniTh, minTh;
    double scaleFactor;            // ORBextractor.h:98 (double member initialised from a float arg)
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> featPerLevel, umax;
    std::vector<Level> lv;
    std::vector<Kp> outKp;
    std::vector<uint8_t> outDesc;

    // ORBextractor.cc:410-470
    Oracle(int nf, float sf, int nl, int ini, int mn) : nfeatures(nf), nlevels(nl), iniTh(ini), minTh(mn), scaleFactor(sf) {
        scale.resize(nl); sigma2.resize(nl); invScale.resize(nl); invSigma2.resize(nl);
        scale[0] = 1.0f; sigma2[0] = 1.0f;
        for (int i = 1; i < nl; i++) {
            scale[i] = (float)(scale[i - 1] * scaleFactor);   // float * double -> float (:421)
            sigma2[i] = scale[i] * scale[i];
        }
        for (int i = 0; i < nl; i++) { invScale[i] = 1.0f / scale[i]; invSigma2[i] = 1.0f / sigma2[i]; }
        lv.resize(nl);
        featPerLevel.resize(nl);
        float factor = (float)(1.0f / scaleFactor);           // :435
        float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
        int sum = 0;
        for (int l = 0; l < nl - 1; l++) {
            featPerLevel[l] = cv_round_f(nDesired);
            sum += featPerLevel[l];
            nDesired *= factor;
        }
        featPerLevel[nl - 1] = std::max(nfeatures - sum, 0);
        // :454-469
        umax.resize(HALF_PATCH_SIZE + 1);
        int v, v0, vmax = (int)floor(HALF_PATCH_SIZE * sqrt(2.f) / 2 + 1);
        int vmin = (int)ceil(HALF_PATCH_SIZE * sqrt(2.f) / 2);
        const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
        for (v = 0; v <= vmax; ++v) umax[v] = cv_round_d(sqrt(hp2 - v * v));
        for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
    }

    // ORBextractor.cc:1107-1132
    void computePyramid(const uint8_t* img, int w, int h, int stride) {
        for (int l = 0; l < nlevels; ++l) {
            Level& L = lv[l];
            L.w = cv_round_f((float)w * invScale[l]);
            L.h = cv_round_f((float)h * invScale[l]);
            L.stride = L.w + 2 * EDGE_THRESHOLD;
            L.buf.assign((size_t)L.stride * (L.h + 2 * EDGE_THRESHOLD), 0);
            if (l != 0) {
                Level& P = lv[l - 1];
                cvl_resize_linear_u8(P.interior(), P.w, P.h, P.stride, L.interior(), L.w, L.h, L.stride);
                cvl_border_reflect101_u8(L.interior(), L.w, L.h, L.stride, L.buf.data(), L.stride, EDGE_THRESHOLD, EDGE_THRESHOLD, EDGE_THRESHOLD, EDGE_THRESHOLD);
            } else {
                cvl_border_reflect101_u8(img, w, h, stride, L.buf.data(), L.stride, EDGE_THRESHOLD, EDGE_THRESHOLD, EDGE_THRESHOLD, EDGE_THRESHOLD);
            }
        }
    }

    // ORBextractor.cc:765-829 (cell loop)
    void detectCells(int level) {
        Level& L = lv[level];
        L.candidates.clear();
        const float W = 30;
        const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
        const int maxBorderX = L.w - EDGE_THRESHOLD + 3, maxBorderY = L.h - EDGE_THRESHOLD + 3;
        const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        const int wCell = (int)ceil(width / nCols), hCell = (int)ceil(height / nRows);
        std::vector<cvl_kp> cell(4096);
        for (int i = 0; i < nRows; i++) {
            const float iniY = (float)(minBorderY + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= maxBorderY - 3) continue;
            if (maxY > maxBorderY) maxY = (float)maxBorderY;
            for (int j = 0; j < nCols; j++) {
                const float iniX = (float)(minBorderX + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= maxBorderX - 6) continue;
                if (maxX > maxBorderX) maxX = (float)maxBorderX;
                const int x0 = (int)iniX, x1 = (int)maxX, y0 = (int)iniY, y1 = (int)maxY;
                const uint8_t* sub = L.interior() + (ptrdiff_t)y0 * L.stride + x0;
                int n = cvl_fast9_16(sub, x1 - x0, y1 - y0, L.stride, iniTh, 1, cell.data(), (int)cell.size());
                if (n == 0) n = cvl_fast9_16(sub, x1 - x0, y1 - y0, L.stride, minTh, 1, cell.data(), (int)cell.size());
                for (int k = 0; k < n; ++k) {
                    Kp kp;
                    kp.x = (float)cell[k].x + j * wCell;       // :822-823
                    kp.y = (float)cell[k].y + i * hCell;
                    kp.size = 7.f; kp.angle = -1.f; kp.response = (float)cell[k].score;
                    kp.octave = 0; kp.class_id = -1;
                    L.candidates.push_back(kp);
                }
            }
        }
    }

    // ExtractorNode::DivideNode, ORBextractor.cc:481-537
    static void divide(const Node& p, const std::vector<Kp>& K, Node c[4]) {
        const int halfX = (int)ceil(static_cast<float>(p.urx - p.ulx) / 2);
        const int halfY = (int)ceil(static_cast<float>(p.bry - p.uly) / 2);
        c[0].ulx = p.ulx; c[0].uly = p.uly; c[0].urx = p.ulx + halfX; c[0].ury = p.uly;
        c[0].blx = p.ulx; c[0].bly = p.uly + halfY; c[0].brx = p.ulx + halfX; c[0].bry = p.uly + halfY;
        c[1].ulx = c[0].urx; c[1].uly = c[0].ury; c[1].urx = p.urx; c[1].ury = p.ury;
        c[1].blx = c[0].brx; c[1].bly = c[0].bry; c[1].brx = p.urx; c[1].bry = p.uly + halfY;
        c[2].ulx = c[0].blx; c[2].uly = c[0].bly; c[2].urx = c[0].brx; c[2].ury = c[0].bry;
        c[2].blx = p.blx; c[2].bly = p.bly; c[2].brx = c[0].brx; c[2].bry = p.bly;
        c[3].ulx = c[2].urx; c[3].uly = c[2].ury; c[3].urx = c[1].brx; c[3].ury = c[1].bry;
        c[3].blx = c[2].brx; c[3].bly = c[2].bry; c[3].brx = p.brx; c[3].bry = p.bry;
        for (size_t i = 0; i < p.keys.size(); i++) {
            const Kp& kp = K[p.keys[i]];
            if (kp.x < c[0].urx) {
                if (kp.y < c[0].bry) c[0].keys.push_back(p.keys[i]);
                else c[2].keys.push_back(p.keys[i]);
            } else if (kp.y < c[0].bry) c[1].keys.push_back(p.keys[i]);
            else c[3].keys.push_back(p.keys[i]);
        }
        for (int k = 0; k < 4; ++k) if (c[k].keys.size() == 1) c[k].noMore = true;
    }

    // ORBextractor.cc:539-763
    std::vector<Kp> distributeOctTree(const std::vector<Kp>& K, int minX, int maxX, int minY, int maxY, int N) {
        std::vector<Kp> result;
        const int nIni = (int)round(static_cast<float>(maxX - minX) / (maxY - minY));
        if (nIni < 1) return result;   // reference: division by zero / UB; rejected here
        const float hX = static_cast<float>(maxX - minX) / nIni;
        std::list<Node> L;
        std::vector<Node*> ini(nIni);
        long counter = 0;
        for (int i = 0; i < nIni; i++) {
            Node n;
            n.ulx = (int)(hX * static_cast<float>(i)); n.uly = 0;
            n.urx = (int)(hX * static_cast<float>(i + 1)); n.ury = 0;
            n.blx = n.ulx; n.bly = maxY - minY;
            n.brx = n.urx; n.bry = maxY - minY;
            n.created = counter++;
            L.push_back(n);
            ini[i] = &L.back();
        }
        for (size_t i = 0; i < K.size(); i++) ini[(size_t)(K[i].x / hX)]->keys.push_back((int)i);
        for (auto it = L.begin(); it != L.end();) {
            if (it->keys.size() == 1) { it->noMore = true; ++it; }
            else if (it->keys.empty()) it = L.erase(it);
            else ++it;
        }
        bool finish = false;
        std::vector<std::pair<int, Node*> > rec;
        auto pushChildren = [&](Node c[4], int& nToExpand) {
            for (int k = 0; k < 4; ++k) {
                if (c[k].keys.empty()) continue;
                c[k].created = counter++;
                L.push_front(c[k]);
                if (c[k].keys.size() > 1) {
                    nToExpand++;
                    rec.push_back(std::make_pair((int)c[k].keys.size(), &L.front()));
                    L.front().self = L.begin();
                }
            }
        };
        while (!finish) {
            int prevSize = (int)L.size();
            auto it = L.begin();
            int nToExpand = 0;
            rec.clear();
            while (it != L.end()) {
                if (it->noMore) { ++it; continue; }
                Node c[4];
                divide(*it, K, c);
                pushChildren(c, nToExpand);
                it = L.erase(it);
            }
            if ((int)L.size() >= N || (int)L.size() == prevSize) finish = true;
            else if (((int)L.size() + nToExpand * 3) > N) {
                while (!finish) {
                    prevSize = (int)L.size();
                    std::vector<std::pair<int, Node*> > prev = rec;
                    rec.clear();
                    // reference :684 sorts (count, pointer); the oracle sorts by count only and keeps
                    // creation order among equals (documented patch)
                    std::stable_sort(prev.begin(), prev.end(), [](const std::pair<int, Node*>& a, const std::pair<int, Node*>& b) { return a.first < b.first; });
                    for (int j = (int)prev.size() - 1; j >= 0; j--) {
                        Node c[4];
                        divide(*prev[j].second, K, c);
                        int dummy = 0;
                        pushChildren(c, dummy);
                        L.erase(prev[j].second->self);
                        if ((int)L.size() >= N) break;
                    }
                    if ((int)L.size() >= N || (int)L.size() == prevSize) finish = true;
                }
            }
        }
        // :742-760
        result.reserve(L.size());
        for (auto it = L.begin(); it != L.end(); ++it) {
            int best = it->keys[0];
            float maxResponse = K[best].response;
            for (size_t k = 1; k < it->keys.size(); k++)
                if (K[it->keys[k]].response > maxResponse) { best = it->keys[k]; maxResponse = K[best].response; }
            result.push_back(K[best]);
        }
        return result;
    }

    // IC_Angle, ORBextractor.cc:77-104
    float icAngle(Level& L, float px, float py) {
        int m_01 = 0, m_10 = 0;
        const uint8_t* center = L.interior() + (ptrdiff_t)cv_round_f(py) * L.stride + cv_round_f(px);
        for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
        const int step = L.stride;
        for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
            int v_sum = 0;
            const int d = umax[v];
            for (int u = -d; u <= d; ++u) {
                int val_plus = center[u + v * step], val_minus = center[u - v * step];
                v_sum += (val_plus - val_minus);
                m_10 += u * (val_plus + val_minus);
            }
            m_01 += v * v_sum;
        }
        return cvl_fast_atan2((float)m_01, (float)m_10);
    }

    // ORBextractor.cc:765-853
    void computeKeyPointsOctTree() {
        for (int level = 0; level < nlevels; ++level) {
            Level& L = lv[level];
            const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
            const int maxBorderX = L.w - EDGE_THRESHOLD + 3, maxBorderY = L.h - EDGE_THRESHOLD + 3;
            detectCells(level);
            L.keypoints = distributeOctTree(L.candidates, minBorderX, maxBorderX, minBorderY, maxBorderY, featPerLevel[level]);
            const int scaledPatchSize = (int)(PATCH_SIZE * scale[level]);
            for (size_t i = 0; i < L.keypoints.size(); i++) {
                L.keypoints[i].x += minBorderX;
                L.keypoints[i].y += minBorderY;
                L.keypoints[i].octave = level;
                L.keypoints[i].size = (float)scaledPatchSize;
            }
        }
        for (int level = 0; level < nlevels; ++level)
            for (Kp& kp : lv[level].keypoints) kp.angle = icAngle(lv[level], kp.x, kp.y);
    }

    // computeOrbDescriptor, ORBextractor.cc:108-147 (img = blurred level, step = its row stride)
    void descriptor(const Kp& kpt, const uint8_t* img, int step, uint8_t* desc) {
        const float factorPI = (float)(CV_PI_D / 180.f);
        float angle = (float)kpt.angle * factorPI;
        float a = cosf(angle), b = sinf(angle);     // `cos(float)` under `using namespace std` is cosf
        const uint8_t* center = img + (ptrdiff_t)cv_round_f(kpt.y) * step + cv_round_f(kpt.x);
        for (int i = 0; i < 32; ++i) {
            int val = 0;
            for (int k = 0; k < 8; ++k) {
                const int i0 = 16 * i + 2 * k, i1 = i0 + 1;
                const int t0 = center[cv_round_f(kPatX[i0] * b + kPatY[i0] * a) * step + cv_round_f(kPatX[i0] * a - kPatY[i0] * b)];
                const int t1 = center[cv_round_f(kPatX[i1] * b + kPatY[i1] * a) * step + cv_round_f(kPatX[i1] * a - kPatY[i1] * b)];
                val |= (t0 < t1) << k;
            }
            desc[i] = (uint8_t)val;
        }
    }
    static constexpr double CV_PI_D = 3.1415926535897932384626433832795;

    // ORBextractor::operator(), ORBextractor.cc:1043-1105
    int extract(const uint8_t* img, int w, int h, int stride) {
        outKp.clear(); outDesc.clear();
        if (!img || w <= 0 || h <= 0) return 0;
        computePyramid(img, w, h, stride);
        computeKeyPointsOctTree();
        int n = 0;
        for (int l = 0; l < nlevels; ++l) n += (int)lv[l].keypoints.size();
        outDesc.assign((size_t)n * 32, 0);
        int offset = 0;
        for (int l = 0; l < nlevels; ++l) {
            Level& L = lv[l];
            L.blurred.clear();
            const int nl = (int)L.keypoints.size();
            if (nl == 0) continue;
            L.blurred.resize((size_t)L.w * L.h);
            cvl_gaussian7x7_u8(L.interior(), L.w, L.h, L.stride, L.blurred.data(), L.w);
            for (int i = 0; i < nl; ++i) descriptor(L.keypoints[i], L.blurred.data(), L.w, outDesc.data() + (size_t)(offset + i) * 32);
            offset += nl;
            for (int i = 0; i < nl; ++i) {
                Kp kp = L.keypoints[i];
                if (l != 0) { kp.x *= scale[l]; kp.y *= scale[l]; }
                outKp.push_back(kp);
            }
        }
        return n;
    }
};

}  // namespace

extern "C" {

void* orbo_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    return new Oracle(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void orbo_destroy(void* h) { delete (Oracle*)h; }

// fills 4*nlevels floats (scale, invScale, sigma2, invSigma2), nlevels ints, 16 ints
void orbo_tables(void* h, float* scales, int* featPerLevel, int* umax) {
    Oracle* o = (Oracle*)h;
    for (int i = 0; i < o->nlevels; ++i) {
        scales[i] = o->scale[i];
        scales[o->nlevels + i] = o->invScale[i];
        scales[2 * o->nlevels + i] = o->sigma2[i];
        scales[3 * o->nlevels + i] = o->invSigma2[i];
        featPerLevel[i] = o->featPerLevel[i];
    }
    for (int i = 0; i < 16; ++i) umax[i] = o->umax[i];
}

int orbo_extract(void* h, const uint8_t* img, int w, int hgt, int stride, void* kp_out, int cap, uint8_t* desc_out) {
    Oracle* o = (Oracle*)h;
    int n = o->extract(img, w, hgt, stride);
    int m = n < cap ? n : cap;
    if (kp_out && m) memcpy(kp_out, o->outKp.data(), (size_t)m * sizeof(Kp));
    if (desc_out && m) memcpy(desc_out, o->outDesc.data(), (size_t)m * 32);
    return n;
}

void orbo_level_dims(void* h, int level, int* w, int* hgt) {
    Oracle* o = (Oracle*)h;
    *w = o->lv[level].w; *hgt = o->lv[level].h;
}
// bordered != 0: copy the whole (w+38)x(h+38) buffer, else the interior; out is tightly packed
void orbo_get_level(void* h, int level, int bordered, uint8_t* out) {
    Oracle* o = (Oracle*)h;
    Level& L = o->lv[level];
    if (bordered) { memcpy(out, L.buf.data(), L.buf.size()); return; }
    for (int y = 0; y < L.h; ++y) memcpy(out + (size_t)y * L.w, L.interior() + (size_t)y * L.stride, L.w);
}
int orbo_get_blurred(void* h, int level, uint8_t* out) {
    Oracle* o = (Oracle*)h;
    Level& L = o->lv[level];
    if (L.blurred.empty()) return 0;
    memcpy(out, L.blurred.data(), L.blurred.size());
    return 1;
}
// which: 0 = FAST candidates (emission order, coords relative to minBorder), 1 = level keypoints
int orbo_get_level_points(void* h, int level, int which, void* out, int cap) {
    Oracle* o = (Oracle*)h;
    const std::vector<Kp>& v = which ? o->lv[level].keypoints : o->lv[level].candidates;
    int m = (int)v.size() < cap ? (int)v.size() : cap;
    if (out && m) memcpy(out, v.data(), (size_t)m * sizeof(Kp));
    return (int)v.size();
}

// Stand-alone stage entry points used by stage-level parity tests --------------------------------
int orbo_octree(const void* cand, int n, int minX, int maxX, int minY, int maxY, int N, void* out, int cap) {
    Oracle o(1000, 1.2f, 8, 20, 7);
    std::vector<Kp> K((const Kp*)cand, (const Kp*)cand + n);
    std::vector<Kp> r = o.distributeOctTree(K, minX, maxX, minY, maxY, N);
    int m = (int)r.size() < cap ? (int)r.size() : cap;
    if (out && m) memcpy(out, r.data(), (size_t)m * sizeof(Kp));
    return (int)r.size();
}

}  // extern "C"
