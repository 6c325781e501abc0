// ORBVocabulary.cc — see ORBVocabulary.h.  Host side of ORBVocabulary::transform on the GPU (C ABI: orbgpu_vocabulary_create,
// orbgpu_bow_transform, include/orbgpu.h).  No distance is computed on the host.
#include "ORBVocabulary.h"

#include <atomic>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>

#include "orbgpu.h"

namespace ORB_SLAM2 {

namespace {
void check(int rc, const char* what) {
    if (rc != 0) throw std::runtime_error(std::string("ORBVocabulary (GPU): ") + what + ": " + orbgpu_last_error());
}

// One vocabulary object is shared by the Tracking, LocalMapping and LoopClosing threads (Tracking.cc:874, LocalMapping.cc:164).
// Each host thread transforms through its own fork of the device handle — same tree in HBM, own stream and scratch — so
// concurrent ComputeBoW calls neither interleave on one stream nor free each other's buffers.  Forks are keyed by the upload's
// generation id (an address could be reused by a later vocabulary) and die with their thread.
std::atomic<unsigned long long> g_next_upload_id{1};
struct ThreadForks {
    std::map<unsigned long long, orbgpu_vocabulary*> forks;
    ~ThreadForks() {
        for (std::map<unsigned long long, orbgpu_vocabulary*>::iterator it = forks.begin(); it != forks.end(); ++it) orbgpu_vocabulary_destroy(it->second);
    }
};
orbgpu_vocabulary* thread_fork(const orbgpu_vocabulary* base, unsigned long long id) {
    thread_local ThreadForks t;
    std::map<unsigned long long, orbgpu_vocabulary*>::iterator it = t.forks.find(id);
    if (it != t.forks.end()) return it->second;
    if (t.forks.size() > 8) {   // vocabularies come and go rarely; keep the per-thread table small
        for (it = t.forks.begin(); it != t.forks.end(); ++it) orbgpu_vocabulary_destroy(it->second);
        t.forks.clear();
    }
    orbgpu_vocabulary* f = 0;
    check(orbgpu_vocabulary_fork(base, &f), "cannot fork the device vocabulary for this thread");
    t.forks[id] = f;
    return f;
}
}  // namespace

ORBVocabulary::ORBVocabulary() : dev_(0), n_words_(0), upload_id_(0) {}

ORBVocabulary::~ORBVocabulary() {
    if (dev_) orbgpu_vocabulary_destroy(dev_);
}

void ORBVocabulary::upload(int k, int L, int scoring, int weighting, const std::vector<int32_t>& parent, const std::vector<uint8_t>& is_leaf,
                           const std::vector<uint8_t>& desc, const std::vector<double>& weight) {
    if (dev_) {
        orbgpu_vocabulary_destroy(dev_);
        dev_ = 0;
    }
    upload_id_ = g_next_upload_id.fetch_add(1);
    check(orbgpu_vocabulary_create(&dev_, orbgpu_default_device(), k, L, scoring, weighting, (int)parent.size(), parent.data(), is_leaf.data(), desc.data(),
                                   weight.data()),
          "cannot upload the vocabulary");
    check(orbgpu_vocabulary_info(dev_, 0, &n_words_), "vocabulary info");
}

#ifdef ORBGPU_SHELL_STANDALONE

// The text format of TemplatedVocabulary::loadFromTextFile (:1338-1423): `k L scoring weighting`, then one line per node,
// `parent isLeaf d0 .. d31 weight`.  Blank lines are skipped (the reference reads an empty last line into uninitialised
// variables, :1378-1395).
bool ORBVocabulary::loadFromTextFile(const std::string& filename) {
    std::ifstream f(filename.c_str());
    if (!f.is_open()) return false;
    std::string s;
    if (!std::getline(f, s)) return false;
    int k = -1, L = -1, n1 = -1, n2 = -1;
    {
        std::stringstream ss(s);
        ss >> k >> L >> n1 >> n2;
    }
    if (k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) return false;   // :1359-1363
    std::vector<int32_t> parent;
    std::vector<uint8_t> is_leaf, desc;
    std::vector<double> weight;
    while (std::getline(f, s)) {
        std::stringstream ss(s);
        int pid, leaf;
        if (!(ss >> pid >> leaf)) continue;
        parent.push_back(pid);
        is_leaf.push_back(leaf > 0 ? 1 : 0);
        for (int i = 0; i < 32; ++i) {
            int b = 0;
            ss >> b;
            desc.push_back((uint8_t)b);
        }
        double w = 0;
        ss >> w;
        weight.push_back(w);
    }
    if (parent.empty()) return false;
    upload(k, L, n1, n2, parent, is_leaf, desc, weight);
    return true;
}

#else

bool ORBVocabulary::loadFromTextFile(const std::string& filename) {
    if (!Base::loadFromTextFile(filename)) return false;
    uploadToDevice();
    return true;
}

// m_nodes (TemplatedVocabulary.h:297-331, :411) -> the records of the C ABI: node i = record i - 1
void ORBVocabulary::uploadToDevice() {
    const size_t n = m_nodes.size();
    if (n < 2) throw std::runtime_error("ORBVocabulary (GPU): the vocabulary is empty");
    std::vector<int32_t> parent(n - 1);
    std::vector<uint8_t> is_leaf(n - 1), desc((n - 1) * 32);
    std::vector<double> weight(n - 1);
    unsigned int next_word = 0;
    for (size_t i = 1; i < n; ++i) {
        const Node& nd = m_nodes[i];
        parent[i - 1] = (int32_t)nd.parent;
        is_leaf[i - 1] = nd.isLeaf() ? 1 : 0;
        std::memcpy(&desc[(i - 1) * 32], nd.descriptor.ptr<unsigned char>(), 32);
        weight[i - 1] = nd.weight;
        // the device numbers the words in node order, as loadFromTextFile (:1407-1414) and create() (:1003-1027) do
        if (nd.isLeaf() && nd.word_id != next_word++)
            throw std::runtime_error("ORBVocabulary (GPU): word ids are not in node order; re-save the vocabulary with saveToTextFile");
    }
    upload(m_k, m_L, (int)m_scoring, (int)m_weighting, parent, is_leaf, desc, weight);
}

#endif

void ORBVocabulary::transform(const std::vector<cv::Mat>& features, DBoW2::BowVector& v, DBoW2::FeatureVector& fv, int levelsup) const {
    v.clear();
    fv.clear();
    if (!dev_) {
#ifndef ORBGPU_SHELL_STANDALONE
        if (Base::empty()) return;   // "safe for subclasses" (:1133-1136)
#endif
        throw std::runtime_error("ORBVocabulary (GPU): the vocabulary has not been uploaded (loadFromTextFile / uploadToDevice)");
    }
    const int n = (int)features.size();
    if (n == 0) return;
    // Converter::toDescriptorVector (Converter.cc:29-37) hands over one 1x32 row header per descriptor: pack them again
    std::vector<uint8_t> rows((size_t)n * 32);
    for (int i = 0; i < n; ++i) std::memcpy(&rows[(size_t)i * 32], features[i].ptr<unsigned char>(), 32);
    const int32_t kp_off[2] = {0, n};
    int32_t bv_off[2], node_off[2];
    std::vector<uint32_t> word(n);
    std::vector<double> value(n);
    std::vector<int32_t> node_id(n), feat_off(n + 1), feat(n);
    check(orbgpu_bow_transform(thread_fork(dev_, upload_id_), 1, kp_off, rows.data(), levelsup, bv_off, word.data(), value.data(), node_off, node_id.data(), feat_off.data(),
                               feat.data(), 0, 0),
          "transform");
    for (int i = 0; i < bv_off[1]; ++i) v.insert(v.end(), std::make_pair((DBoW2::WordId)word[i], (DBoW2::WordValue)value[i]));
    for (int j = 0; j < node_off[1]; ++j) {
        std::vector<unsigned int>& dst = fv.insert(fv.end(), std::make_pair((DBoW2::NodeId)node_id[j], std::vector<unsigned int>()))->second;
        dst.assign(feat.begin() + feat_off[j], feat.begin() + feat_off[j + 1]);
    }
}

}  // namespace ORB_SLAM2
