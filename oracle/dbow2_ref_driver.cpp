// dbow2_ref_driver.cpp -- C entry points around the REFERENCE'S OWN DBoW2 (compiled from /root/reference/src/sg-slam/Thirdparty/DBoW2 where it
// lies, against the cv::Mat stand-in in oracle/dbow2_shim/): the pin of the bag-of-words restatement (oracle/bow.cpp) and of the vocabulary
// file readers.  TEST INFRASTRUCTURE -- built into oracle/_ref/libdbow2_ref.so by oracle/Makefile, loaded only by tests/.
// The vocabulary type is the one ORB-SLAM2 uses (include/ORBVocabulary.h:31: TemplatedVocabulary<FORB::TDescriptor, FORB>).
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "DBoW2/FORB.h"
#include "DBoW2/TemplatedVocabulary.h"

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {
typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> OrbVoc;
struct Voc : OrbVoc {                       // derived only to read the protected node table
    int n_nodes() const { return (int)m_nodes.size(); }
    void dump(int32_t* parent, uint8_t* desc, double* weight, int32_t* word_id, int32_t* nchildren) const {
        for (size_t i = 0; i < m_nodes.size(); ++i) {
            const Node& nd = m_nodes[i];
            parent[i] = i == 0 ? -1 : (int32_t)nd.parent;
            weight[i] = nd.weight;
            word_id[i] = nd.isLeaf() && i != 0 ? (int32_t)nd.word_id : -1;
            nchildren[i] = (int32_t)nd.children.size();
            if (i != 0 && !nd.descriptor.empty()) std::memcpy(desc + 32 * i, nd.descriptor.ptr<unsigned char>(), 32);
            else std::memset(desc + 32 * i, 0, 32);
        }
    }
    int child(int node, int j) const { return (int)m_nodes[node].children[j]; }
    void each(const cv::Mat& f, DBoW2::WordId& id, DBoW2::WordValue& w, DBoW2::NodeId* nid, int levelsup) const { transform(f, id, w, nid, levelsup); }
};
cv::Mat row_of(const uint8_t* d) { cv::Mat m(1, 32, CV_8U); std::memcpy(m.ptr<unsigned char>(), d, 32); return m; }
}  // namespace

REF_API void* dbow2_ref_create(const uint8_t* desc, const int32_t* image_of, int n, int nimages, int k, int L) {
    std::vector<std::vector<cv::Mat>> feats(nimages);
    for (int i = 0; i < n; ++i) feats[image_of[i]].push_back(row_of(desc + 32 * (size_t)i));
    Voc* v = new Voc();
    v->create(feats, k, L, DBoW2::TF_IDF, DBoW2::L1_NORM);
    return v;
}
REF_API void dbow2_ref_free(void* h) { delete static_cast<Voc*>(h); }
REF_API void* dbow2_ref_load_text(const char* path) { Voc* v = new Voc(); if (!v->loadFromTextFile(path)) { delete v; return nullptr; } return v; }
REF_API void* dbow2_ref_load_binary(const char* path) { Voc* v = new Voc(); if (!v->loadFromBinaryFile(path)) { delete v; return nullptr; } return v; }
REF_API void dbow2_ref_save_text(void* h, const char* path) { static_cast<Voc*>(h)->saveToTextFile(path); }
REF_API void dbow2_ref_save_binary(void* h, const char* path) { static_cast<Voc*>(h)->saveToBinaryFile(path); }
REF_API int dbow2_ref_info(void* h, int* k, int* L, int* nwords) {
    Voc* v = static_cast<Voc*>(h);
    *k = v->getBranchingFactor(); *L = v->getDepthLevels(); *nwords = (int)v->size();
    return v->n_nodes();
}
REF_API void dbow2_ref_dump(void* h, int32_t* parent, uint8_t* desc, double* weight, int32_t* word_id, int32_t* nchildren) {
    static_cast<Voc*>(h)->dump(parent, desc, weight, word_id, nchildren);
}
REF_API int dbow2_ref_child(void* h, int node, int j) { return static_cast<Voc*>(h)->child(node, j); }

// Frame::ComputeBoW (src/Frame.cc:421-428): transform(vCurrentDesc, mBowVec, mFeatVec, 4).  BowVector as (ascending word id, value) pairs,
// FeatureVector flattened in std::map order as (node id, feature index) pairs.  Returns the number of BowVector entries; *nfv = feature-vector pairs.
REF_API int dbow2_ref_transform(void* h, const uint8_t* desc, int n, int levelsup, int32_t* bow_word, double* bow_value, int32_t* fv_node, int32_t* fv_feature, int* nfv) {
    std::vector<cv::Mat> f;
    for (int i = 0; i < n; ++i) f.push_back(row_of(desc + 32 * (size_t)i));
    DBoW2::BowVector bv; DBoW2::FeatureVector fv;
    static_cast<Voc*>(h)->transform(f, bv, fv, levelsup);
    int c = 0;
    for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++c) { bow_word[c] = (int32_t)it->first; bow_value[c] = it->second; }
    int q = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it)
        for (size_t j = 0; j < it->second.size(); ++j, ++q) { fv_node[q] = (int32_t)it->first; fv_feature[q] = (int32_t)it->second[j]; }
    *nfv = q;
    return c;
}
// the per-feature form (word id, weight, node id levelsup levels above the leaf)
REF_API void dbow2_ref_transform_each(void* h, const uint8_t* desc, int n, int levelsup, int32_t* word, double* weight, int32_t* node) {
    Voc* v = static_cast<Voc*>(h);
    for (int i = 0; i < n; ++i) {
        DBoW2::WordId id; DBoW2::WordValue w; DBoW2::NodeId nid = 0;
        v->each(row_of(desc + 32 * (size_t)i), id, w, &nid, levelsup);
        word[i] = (int32_t)id; weight[i] = w; node[i] = (int32_t)nid;
    }
}
// FORB::distance (Thirdparty/DBoW2/DBoW2/FORB.cpp:81-101) == ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1649-1665)
REF_API int dbow2_ref_distance(const uint8_t* a, const uint8_t* b) { return DBoW2::FORB::distance(row_of(a), row_of(b)); }
REF_API double dbow2_ref_score(void* h, const int32_t* w1, const double* v1, int n1, const int32_t* w2, const double* v2, int n2) {
    DBoW2::BowVector a, b;
    for (int i = 0; i < n1; ++i) a.addWeight(w1[i], v1[i]);
    for (int i = 0; i < n2; ++i) b.addWeight(w2[i], v2[i]);
    return static_cast<Voc*>(h)->score(a, b);
}
