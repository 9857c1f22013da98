// ORACLE (test infrastructure, NOT product code): C entry points around the reference's vendored DBoW2
// (orb_slam3/Thirdparty/DBoW2, compiled unmodified against cvshim/): the ORBVocabulary typedef of ORBVocabulary.h:28-29,
// loaded with loadFromTextFile (System.cc:116) and used as Frame::ComputeBoW does (Frame.cc:738-745).
#include <cstdint>
#include <cstring>

#include "DBoW2/FORB.h"
#include "DBoW2/TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;

extern "C" {

void* refbow_load_text(const char* path) {
    ORBVocabulary* v = new ORBVocabulary();
    if (!v->loadFromTextFile(path)) { delete v; return nullptr; }
    return v;
}
void refbow_destroy(void* v) { delete (ORBVocabulary*)v; }
int refbow_size(void* v) { return (int)((ORBVocabulary*)v)->size(); }

int refbow_distance(const uint8_t* a, const uint8_t* b) {                     // FORB::distance == ORBmatcher::DescriptorDistance's bit trick
    cv::Mat ma(1, 32, CV_8U), mb(1, 32, CV_8U);                                // (own, 4-byte aligned storage: the function reads int32_t words)
    memcpy(ma.data, a, 32); memcpy(mb.data, b, 32);
    return DBoW2::FORB::distance(ma, mb);
}

// Frame::ComputeBoW: Converter::toDescriptorVector (one 1x32 Mat per row, Converter.cc:26-34) then transform(.., levelsup).
// Outputs in std::map order: words {id, value}; feature-vector nodes {node, start into fvFeat}; counts = {nWords, nNodes, nFeat}.
int refbow_transform(void* vp, const uint8_t* desc, int n, int levelsup, int32_t* bowId, double* bowVal, int32_t* fvNode,
                     int32_t* fvStart, int32_t* fvFeat, int32_t* counts) {
    const ORBVocabulary* voc = (const ORBVocabulary*)vp;
    std::vector<cv::Mat> features;
    features.reserve(n);
    for (int i = 0; i < n; i++) {
        cv::Mat row(1, 32, CV_8U);
        memcpy(row.data, desc + (size_t)32 * i, 32);
        features.push_back(row);
    }
    DBoW2::BowVector bow;
    DBoW2::FeatureVector fv;
    voc->transform(features, bow, fv, levelsup);
    int w = 0;
    for (DBoW2::BowVector::const_iterator it = bow.begin(); it != bow.end(); ++it, ++w) { bowId[w] = (int32_t)it->first; bowVal[w] = it->second; }
    int g = 0, f = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++g) {
        fvNode[g] = (int32_t)it->first; fvStart[g] = f;
        for (size_t k = 0; k < it->second.size(); k++) fvFeat[f++] = (int32_t)it->second[k];
    }
    counts[0] = w; counts[1] = g; counts[2] = f;
    return 0;
}

}  // extern "C"
