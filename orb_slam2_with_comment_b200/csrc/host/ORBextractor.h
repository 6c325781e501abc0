// ORBextractor.h — drop-in replacement of ORB_SLAM2::ORBextractor (/root/reference/include/ORBextractor.h:45-111).
//
// Same namespace, class name, constructor, operator(), getters and public mvImagePyramid, so Frame.cc, Tracking.cc
// and System.cc compile and link against it unchanged.  All pixel work happens in liborbgpu.so (CUDA, sm_100a)
// through the C ABI of include/orbgpu.h; there is no CPU path: a failing GPU call throws std::runtime_error.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <list>
#include <vector>

#include <opencv/cv.h>

struct orbgpu_extractor;

namespace ORB_SLAM2 {

// Kept for source compatibility only (declared in the reference header, :32-43); the quadtree runs on the device.
class ExtractorNode {
public:
    ExtractorNode() : bNoMore(false) {}
    std::vector<cv::KeyPoint> vKeys;
    cv::Point2i UL, UR, BL, BR;
    std::list<ExtractorNode>::iterator lit;
    bool bNoMore;
};

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image; the mask is ignored, as in the reference (:58).
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors);

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // Filled after every call like the reference's (ComputePyramid, :1107-1132): level l is a w x h view inside a
    // buffer that carries the 19-px BORDER_REFLECT_101 frame.  Frame::ComputeStereoMatches reads it
    // (Frame.cc:508,598,610,615).  Monocular callers that never look at it can switch the copy off.
    std::vector<cv::Mat> mvImagePyramid;
    void SetPyramidDownload(bool on) { mbDownloadPyramid = on; }

    // Frame::ComputeStereoMatches (Frame.cc:501-675) on the device: pLeft / pRight are the two extractors that have just
    // processed the left and right image of the frame (Frame.cc:78-83); fills mvuRight / mvDepth for the N left key points.
    // In Frame.cc the body of ComputeStereoMatches() becomes
    //     ORBextractor::ComputeStereoMatches(mpORBextractorLeft, mpORBextractorRight, N, mb, mbf, mvuRight, mvDepth);
    // and the pyramids never leave the GPU (SetPyramidDownload(false) on both).
    static void ComputeStereoMatches(ORBextractor* pLeft, ORBextractor* pRight, int N, float mb, float mbf,
                                     std::vector<float>& mvuRight, std::vector<float>& mvDepth);

    // Device placement of instances created afterwards (default: device 0).
    static void SetDevice(int device);

protected:
    void EnsureHandle(int width, int height);

    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;

    std::vector<int> mnFeaturesPerLevel;
    std::vector<int> umax;
    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

    orbgpu_extractor* mpHandle;
    int mnMaxWidth, mnMaxHeight;
    bool mbDownloadPyramid;
    std::vector<cv::KeyPoint> mvKeyBuffer;   // kp_capacity records, reused between calls
    std::vector<unsigned char> mvDescBuffer;
    std::vector<cv::Mat> mvBordered;         // owners of the bordered level buffers behind mvImagePyramid
};

}  // namespace ORB_SLAM2

#endif
