// ORBextractor.cc — host shell of the drop-in ORBextractor: argument checks, buffer ownership and the calls into the
// C ABI.  No pixel is touched here.
#include "ORBextractor.h"

#include <cassert>
#include <cstring>
#include <stdexcept>
#include <string>

#include "orbgpu.h"

namespace ORB_SLAM2 {

static_assert(sizeof(cv::KeyPoint) == sizeof(orbgpu_keypoint), "cv::KeyPoint must be the 28-byte record the C ABI writes");

namespace {
void check(int rc, const char* what) {
    if (rc != 0) throw std::runtime_error(std::string("ORBextractor (GPU): ") + what + ": " + orbgpu_last_error());
}
}  // namespace

void ORBextractor::SetDevice(int device) { check(orbgpu_set_default_device(device), "SetDevice"); }

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST),
      mpHandle(nullptr), mnMaxWidth(0), mnMaxHeight(0), mbDownloadPyramid(true) {
    // the constructor's tables (reference :413-469) come from the library so that both sides share one derivation
    std::vector<float> t(4 * (size_t)nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    umax.resize(16);
    check(orbgpu_extractor_static_tables(nfeatures, _scaleFactor, nlevels, t.data(), mnFeaturesPerLevel.data(), umax.data()),
          "bad constructor arguments");
    mvScaleFactor.assign(t.begin(), t.begin() + nlevels);
    mvInvScaleFactor.assign(t.begin() + nlevels, t.begin() + 2 * nlevels);
    mvLevelSigma2.assign(t.begin() + 2 * nlevels, t.begin() + 3 * nlevels);
    mvInvLevelSigma2.assign(t.begin() + 3 * nlevels, t.end());
    mvImagePyramid.resize(nlevels);
}

ORBextractor::~ORBextractor() {
    if (mpHandle) orbgpu_extractor_destroy(mpHandle);
}

// The workspace is sized by the first image; a larger image later re-creates it.
void ORBextractor::EnsureHandle(int width, int height) {
    if (mpHandle && width <= mnMaxWidth && height <= mnMaxHeight) return;
    if (mpHandle) {
        orbgpu_extractor_destroy(mpHandle);
        mpHandle = nullptr;
    }
    mnMaxWidth = width > mnMaxWidth ? width : mnMaxWidth;
    mnMaxHeight = height > mnMaxHeight ? height : mnMaxHeight;
    check(orbgpu_extractor_create(&mpHandle, orbgpu_default_device(), nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST, mnMaxWidth,
                                  mnMaxHeight, 1),
          "cannot create the device extractor");
    const int cap = orbgpu_extractor_max_keypoints(mpHandle);
    mvKeyBuffer.resize(cap);
    mvDescBuffer.resize((size_t)cap * 32);
}

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints,
                              cv::OutputArray _descriptors) {
    if (_image.empty()) return;   // reference :1046-1047
    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1);
    EnsureHandle(image.cols, image.rows);

    int n = 0;
    check(orbgpu_extract(mpHandle, image.ptr(0), image.cols, image.rows, (size_t)image.step,
                         reinterpret_cast<orbgpu_keypoint*>(mvKeyBuffer.data()), mvDescBuffer.data(), (int)mvKeyBuffer.size(), &n),
          "extraction failed");

    _keypoints.assign(mvKeyBuffer.begin(), mvKeyBuffer.begin() + n);
    if (n == 0) {
        _descriptors.release();   // :1064-1065
    } else {
        _descriptors.create(n, 32, CV_8U);
        cv::Mat descriptors = _descriptors.getMat();
        for (int i = 0; i < n; ++i) std::memcpy(descriptors.ptr(i), mvDescBuffer.data() + (size_t)i * 32, 32);
    }

    if (mbDownloadPyramid) {
        const int E = 19;   // EDGE_THRESHOLD
        mvBordered.resize(nlevels);
        for (int l = 0; l < nlevels; ++l) {
            int w = 0, h = 0;
            check(orbgpu_extractor_level_dims(mpHandle, l, &w, &h), "level size");
            mvBordered[l].create(h + 2 * E, w + 2 * E, CV_8UC1);
            check(orbgpu_extractor_read_level(mpHandle, 0, l, 1, mvBordered[l].ptr(0), (size_t)mvBordered[l].step), "pyramid download");
            mvImagePyramid[l] = mvBordered[l](cv::Rect(E, E, w, h));
        }
    }
}

void ORBextractor::ComputeStereoMatches(ORBextractor* pLeft, ORBextractor* pRight, int N, float mb, float mbf,
                                        std::vector<float>& mvuRight, std::vector<float>& mvDepth) {
    mvuRight = std::vector<float>(N, -1.0f);   // Frame.cc:503-504
    mvDepth = std::vector<float>(N, -1.0f);
    if (N == 0 || !pLeft->mpHandle || !pRight->mpHandle) return;
    const int cap = orbgpu_extractor_max_keypoints(pLeft->mpHandle);
    std::vector<float> u(cap, -1.0f), d(cap, -1.0f);
    check(orbgpu_stereo_matches(pLeft->mpHandle, pRight->mpHandle, mb, mbf, u.data(), d.data(), cap), "stereo matching failed");
    for (int i = 0; i < N && i < cap; ++i) { mvuRight[i] = u[i]; mvDepth[i] = d[i]; }
}

}  // namespace ORB_SLAM2
