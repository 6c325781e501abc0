// vocab_derived_test.cc — TEST INFRASTRUCTURE.  The ORBVocabulary shell in its deployment form (derived from the
// reference's DBoW2::TemplatedVocabulary, csrc/host/ORBVocabulary.h without ORBGPU_SHELL_STANDALONE), linked with the
// reference's own DBoW2 sources: the SAME object answers transform() twice — through the GPU override and through the
// reference's CPU implementation (qualified call) — and the two BowVectors / FeatureVectors must be identical, doubles
// included.  Built only where /root/reference exists (oracle/Makefile -> oracle/_ref/vocab_derived_test); the binary travels.
//   usage: vocab_derived_test <vocabulary.txt> <descriptors.bin (n x 32 bytes)> <levelsup>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <vector>

#include "ORBVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> RefVocabulary;

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    ORB_SLAM2::ORBVocabulary voc;
    try {
        if (!voc.loadFromTextFile(argv[1])) { printf("load failed\n"); return 2; }
    } catch (const std::runtime_error& e) {
        printf("%s\n", e.what());
        return 3;   // no device: fails loudly
    }
    FILE* f = fopen(argv[2], "rb");
    if (!f) return 2;
    std::vector<unsigned char> bytes;
    unsigned char buf[4096];
    size_t got;
    while ((got = fread(buf, 1, sizeof buf, f)) > 0) bytes.insert(bytes.end(), buf, buf + got);
    fclose(f);
    const int n = (int)(bytes.size() / 32), levelsup = atoi(argv[3]);
    cv::Mat D(n, 32, CV_8U, bytes.data());
    std::vector<cv::Mat> feats;
    for (int i = 0; i < n; ++i) feats.push_back(D.row(i));
    DBoW2::BowVector bv_gpu, bv_ref;
    DBoW2::FeatureVector fv_gpu, fv_ref;
    const RefVocabulary& as_base = voc;
    as_base.transform(feats, bv_gpu, fv_gpu, levelsup);               // virtual -> the GPU override, as Frame::ComputeBoW reaches it
    voc.RefVocabulary::transform(feats, bv_ref, fv_ref, levelsup);    // the reference's CPU code
    int bad = bv_gpu.size() != bv_ref.size() || fv_gpu.size() != fv_ref.size();
    DBoW2::BowVector::const_iterator a = bv_gpu.begin(), b = bv_ref.begin();
    for (; a != bv_gpu.end() && b != bv_ref.end(); ++a, ++b) bad += a->first != b->first || a->second != b->second;
    DBoW2::FeatureVector::const_iterator c = fv_gpu.begin(), d = fv_ref.begin();
    for (; c != fv_gpu.end() && d != fv_ref.end(); ++c, ++d) bad += c->first != d->first || c->second != d->second;
    const double s = voc.score(bv_gpu, bv_ref);   // the reference's L1 score of a vector with itself
    printf("vocab_derived_test: %d features, %zu words, %zu nodes, score(gpu, ref) = %.17g, %d differences\n", n, bv_gpu.size(), fv_gpu.size(), s, bad);
    return bad ? 1 : 0;
}
