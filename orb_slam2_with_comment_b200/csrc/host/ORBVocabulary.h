// ORBVocabulary.h — drop-in for the reference's include/ORBVocabulary.h (:31-32, a typedef of
// DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB>): the same class with the one operation on the hot
// path moved to the GPU,
//
//     transform(const std::vector<cv::Mat>& features, BowVector&, FeatureVector&, int levelsup)    TemplatedVocabulary.h:1127-1197
//
// which is what Frame::ComputeBoW (Frame.cc:425-432) and KeyFrame::ComputeBoW (KeyFrame.cc:59-70) call.  In an ORB-SLAM2
// build ORBVocabulary DERIVES from the reference's template (transform is virtual there, :145), so loading, score(), the
// other transform overloads, size() ... are the reference's own code and System.cc:68 / KeyFrameDatabase / LoopClosing
// compile unchanged; loadFromTextFile additionally uploads the tree to the device.  Built with ORBGPU_SHELL_STANDALONE
// (this repository's tests: no DBoW2, no OpenCV) the class stands alone with its own text loader.
// There is no CPU fallback: without a CUDA device loadFromTextFile / transform throw std::runtime_error.
#ifndef ORBGPU_HOST_ORBVOCABULARY_H
#define ORBGPU_HOST_ORBVOCABULARY_H

#include <string>
#include <vector>

#include <opencv2/core/core.hpp>

#ifdef ORBGPU_SHELL_STANDALONE
#include "Frame.h"   // stand-ins of DBoW2::BowVector / FeatureVector (shim/orbslam2)
#else
#include "Thirdparty/DBoW2/DBoW2/FORB.h"
#include "Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"
#endif

struct orbgpu_vocabulary;

namespace ORB_SLAM2 {

#ifdef ORBGPU_SHELL_STANDALONE
class ORBVocabulary {
#else
class ORBVocabulary : public DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> {
    typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> Base;
#endif
public:
    ORBVocabulary();
    virtual ~ORBVocabulary();

    // TemplatedVocabulary::loadFromTextFile (:1338-1423) + upload of the tree to the device
    bool loadFromTextFile(const std::string& filename);

#ifdef ORBGPU_SHELL_STANDALONE
    virtual void transform(const std::vector<cv::Mat>& features, DBoW2::BowVector& v, DBoW2::FeatureVector& fv, int levelsup) const;
    unsigned int size() const { return (unsigned int)n_words_; }
    bool empty() const { return n_words_ == 0; }
#else
    using Base::transform;
    virtual void transform(const std::vector<cv::Mat>& features, DBoW2::BowVector& v, DBoW2::FeatureVector& fv, int levelsup) const;
    // call after any other way of filling the vocabulary (create(), load(), loadFromBinaryFile ...)
    void uploadToDevice();
#endif

private:
    ORBVocabulary(const ORBVocabulary&);
    ORBVocabulary& operator=(const ORBVocabulary&);
    void upload(int k, int L, int scoring, int weighting, const std::vector<int32_t>& parent, const std::vector<uint8_t>& is_leaf,
                const std::vector<uint8_t>& desc, const std::vector<double>& weight);
    orbgpu_vocabulary* dev_;
    int n_words_;
    unsigned long long upload_id_;   // generation of the device copy: keys the per-thread forks (ORBVocabulary.cc)
};

}  // namespace ORB_SLAM2
#endif
