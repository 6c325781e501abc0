// TEST INFRASTRUCTURE — not part of the product path.
// Thin C wrapper around the REFERENCE's own DBoW2 (Thirdparty/DBoW2/DBoW2/{TemplatedVocabulary.h, FORB.cpp, BowVector.cpp,
// FeatureVector.cpp, ScoringObject.cpp} + DUtils/Random.cpp), compiled where the sources lie under /root/reference by
// oracle/Makefile (target dbowref -> oracle/_ref/libdbowref.so) against the cv:: stand-in.  It exposes ORBVocabulary
// (include/ORBVocabulary.h:31-32) loadFromTextFile + transform(features, BowVector, FeatureVector, levelsup), i.e. exactly
// what Frame::ComputeBoW (Frame.cc:425-432) and KeyFrame::ComputeBoW (KeyFrame.cc:59-70) call, with flat outputs.
#include <cstdint>
#include <string>
#include <vector>

#include "FORB.h"
#include "TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;

extern "C" {

void* dbowref_load(const char* text_path) {
    ORBVocabulary* v = new ORBVocabulary();
    if (!v->loadFromTextFile(text_path) || v->empty()) {
        delete v;
        return nullptr;
    }
    return v;
}

void dbowref_free(void* h) { delete static_cast<ORBVocabulary*>(h); }

int dbowref_words(void* h) { return (int)static_cast<ORBVocabulary*>(h)->size(); }

// desc: n rows of 32 bytes.  Outputs sized for n entries (fv_off: n+1).  Returns 0.
int dbowref_transform(void* h, const uint8_t* desc, int n, int levelsup, int* bv_n, uint32_t* bv_word, double* bv_val, int* fv_n,
                      uint32_t* fv_node, int32_t* fv_off, uint32_t* fv_feat) {
    const ORBVocabulary* voc = static_cast<const ORBVocabulary*>(h);
    std::vector<cv::Mat> feats;   // Converter::toDescriptorVector (Converter.cc:29-37): one 1x32 row header per descriptor
    feats.reserve(n);
    cv::Mat all(n, 32, CV_8U, const_cast<uint8_t*>(desc));
    for (int i = 0; i < n; ++i) feats.push_back(all.row(i));
    DBoW2::BowVector bv;
    DBoW2::FeatureVector fv;
    voc->transform(feats, bv, fv, levelsup);
    int k = 0;
    for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++k) {
        bv_word[k] = it->first;
        bv_val[k] = it->second;
    }
    *bv_n = k;
    int nn = 0, pos = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++nn) {
        fv_node[nn] = it->first;
        fv_off[nn] = pos;
        for (size_t j = 0; j < it->second.size(); ++j) fv_feat[pos++] = it->second[j];
    }
    fv_off[nn] = pos;
    *fv_n = nn;
    return 0;
}

double dbowref_score(void* h, int n1, const uint32_t* w1, const double* v1, int n2, const uint32_t* w2, const double* v2) {
    DBoW2::BowVector a, b;
    for (int i = 0; i < n1; ++i) a.insert(a.end(), DBoW2::BowVector::value_type(w1[i], v1[i]));
    for (int i = 0; i < n2; ++i) b.insert(b.end(), DBoW2::BowVector::value_type(w2[i], v2[i]));
    return static_cast<const ORBVocabulary*>(h)->score(a, b);
}

}  // extern "C"
