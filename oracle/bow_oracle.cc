// TEST INFRASTRUCTURE — not part of the product path (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
// may use anything under oracle/).
//
// CPU restatement of the DBoW2 vocabulary descent that produces BowVector / FeatureVector for a frame
// (Frame::ComputeBoW, Frame.cc:425-432; KeyFrame::ComputeBoW, KeyFrame.cc:59-70), over flat arrays:
//   * vocabulary as loadFromTextFile reads it            TemplatedVocabulary.h:1338-1423
//   * transform(features, BowVector, FeatureVector, L')  TemplatedVocabulary.h:1127-1197
//   * transform(feature, word, weight, nid, levelsup)    TemplatedVocabulary.h:1214-1259
//   * FORB::distance                                     FORB.cpp:81-101
//   * BowVector::addWeight / addIfNotExist / normalize   BowVector.cpp:36-90
//   * FeatureVector::addFeature                          FeatureVector.cpp:30-45
//   * mustNormalize per scoring type                     ScoringObject.h:74-89
// Pinned against the reference's own DBoW2 compiled from its sources (oracle/dbow_ref.cc -> oracle/_ref/libdbowref.so) by
// tests/test_oracle_vocabulary.py and the fixtures in tests/golden/vocabulary_golden.npz.
//
// One case where the reference has no defined result: a leaf shallower than level L - levelsup leaves `nid` uninitialised
// (TemplatedVocabulary.h:1156, :1250-1251).  Here (and in the CUDA path) the leaf's own node id is reported.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <thread>
#include <vector>

namespace {

struct Voc {
    int k, L, scoring, weighting;
    std::vector<int32_t> parent;               // per node (node 0 = root)
    std::vector<std::vector<int32_t>> children;
    std::vector<uint8_t> desc;                 // 32 B per node
    std::vector<double> weight;
    std::vector<int32_t> word_id;              // -1 for inner nodes
    int n_words;
};

int forb_distance(const uint8_t* a, const uint8_t* b) {
    int32_t pa[8], pb[8];
    std::memcpy(pa, a, 32);
    std::memcpy(pb, b, 32);
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        unsigned int v = (unsigned int)(pa[i] ^ pb[i]);
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

void descend(const Voc& V, const uint8_t* f, int levelsup, uint32_t* word, double* w, uint32_t* nid) {
    const int nid_level = V.L - levelsup;
    bool have = false;
    if (nid_level <= 0) { *nid = 0; have = true; }
    int32_t cur = 0;
    int level = 0;
    do {
        ++level;
        const std::vector<int32_t>& ch = V.children[cur];
        cur = ch[0];
        double best = forb_distance(f, &V.desc[(size_t)cur * 32]);
        for (size_t j = 1; j < ch.size(); ++j) {
            const double d = forb_distance(f, &V.desc[(size_t)ch[j] * 32]);
            if (d < best) { best = d; cur = ch[j]; }
        }
        if (level == nid_level) { *nid = (uint32_t)cur; have = true; }
    } while (!V.children[cur].empty());
    if (!have) *nid = (uint32_t)cur;
    *word = (uint32_t)V.word_id[cur];
    *w = V.weight[cur];
}

}  // namespace

extern "C" {

// records = the node lines of the text file in order; node id = 1 + record index.
void* orbo_voc_create(int k, int L, int scoring, int weighting, int n_records, const int32_t* parent, const uint8_t* is_leaf,
                      const uint8_t* desc, const double* weight) {
    Voc* V = new Voc();
    V->k = k; V->L = L; V->scoring = scoring; V->weighting = weighting;
    const int n = n_records + 1;
    V->parent.assign(n, 0);
    V->children.assign(n, std::vector<int32_t>());
    V->desc.assign((size_t)n * 32, 0);
    V->weight.assign(n, 0.0);
    V->word_id.assign(n, -1);
    V->n_words = 0;
    for (int r = 0; r < n_records; ++r) {
        const int nid = r + 1, pid = parent[r];
        if (pid < 0 || pid >= nid) { delete V; return nullptr; }
        V->parent[nid] = pid;
        V->children[pid].push_back(nid);
        std::memcpy(&V->desc[(size_t)nid * 32], desc + (size_t)r * 32, 32);
        V->weight[nid] = weight[r];
        if (is_leaf[r]) V->word_id[nid] = V->n_words++;
    }
    return V;
}

void orbo_voc_free(void* h) { delete static_cast<Voc*>(h); }
int orbo_voc_words(void* h) { return static_cast<Voc*>(h)->n_words; }

int orbo_voc_transform(void* h, const uint8_t* desc, int n, int levelsup, int* bv_n, uint32_t* bv_word, double* bv_val, int* fv_n,
                       uint32_t* fv_node, int32_t* fv_off, uint32_t* fv_feat, uint32_t* feat_word, uint32_t* feat_node) {
    const Voc& V = *static_cast<Voc*>(h);
    std::map<uint32_t, double> bv;
    std::map<uint32_t, std::vector<uint32_t>> fv;
    const bool accumulate = V.weighting == 0 || V.weighting == 1;   // TF_IDF, TF
    const bool must = V.scoring != 5;                               // all but DOT_PRODUCT
    const bool l2 = V.scoring == 1;
    for (int i = 0; i < n; ++i) {
        uint32_t word, nid;
        double w;
        descend(V, desc + (size_t)i * 32, levelsup, &word, &w, &nid);
        if (feat_word) feat_word[i] = word;
        if (feat_node) feat_node[i] = nid;
        if (w > 0) {
            std::map<uint32_t, double>::iterator it = bv.find(word);
            if (it == bv.end()) bv[word] = w;
            else if (accumulate) it->second += w;
            fv[nid].push_back((uint32_t)i);
        }
    }
    if (accumulate && !bv.empty() && !must) {
        const double nd = (double)bv.size();
        for (auto& e : bv) e.second /= nd;
    }
    if (must) {
        double norm = 0.0;
        if (!l2) {
            for (auto& e : bv) norm += std::fabs(e.second);
        } else {
            for (auto& e : bv) norm += e.second * e.second;
            norm = std::sqrt(norm);
        }
        if (norm > 0.0)
            for (auto& e : bv) e.second /= norm;
    }
    int k = 0;
    for (auto& e : bv) { bv_word[k] = e.first; bv_val[k] = e.second; ++k; }
    *bv_n = k;
    int nn = 0, pos = 0;
    for (auto& e : fv) {
        fv_node[nn] = e.first;
        fv_off[nn] = pos;
        for (uint32_t j : e.second) fv_feat[pos++] = j;
        ++nn;
    }
    fv_off[nn] = pos;
    *fv_n = nn;
    return 0;
}

// Timed CPU leg of bench.py: n_frames frames of `per` descriptors each, one frame per thread slot; returns nothing.
void orbo_voc_bench(void* h, const uint8_t* desc, int n_frames, int per, int levelsup, int n_threads) {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back([=]() {
            std::vector<uint32_t> bw(per), fn(per), ff(per);
            std::vector<double> bvv(per);
            std::vector<int32_t> fo(per + 1);
            int a, b;
            for (int f = t; f < n_frames; f += n_threads)
                orbo_voc_transform(h, desc + (size_t)f * per * 32, per, levelsup, &a, bw.data(), bvv.data(), &b, fn.data(), fo.data(), ff.data(),
                                   nullptr, nullptr);
        });
    for (auto& t : th) t.join();
}

}  // extern "C"
