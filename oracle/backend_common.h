// TEST INFRASTRUCTURE -- pieces shared by the restated and the verbatim SD-tree backends.
#pragma once
#include <algorithm>
#include <limits>
#include "../include/ppg.h"

namespace ppgo {

// "Distribution statistics", GP:1121-1186 -- generic over the accessor interface
template <class B> static inline void backend_statistics(const B &b, ppg_iteration_stats &st) {
    int maxDepth = 0, minDepth = std::numeric_limits<int>::max(); float avgDepth = 0;
    float maxAvgRadiance = 0, minAvgRadiance = std::numeric_limits<float>::max(), avgAvgRadiance = 0;
    size_t maxNodes = 0, minNodes = std::numeric_limits<size_t>::max(); float avgNodes = 0;
    float maxW = 0, minW = std::numeric_limits<float>::max(), avgW = 0;
    int nPoints = 0, nPointsNodes = 0; uint32_t leaves = 0;
    for (size_t i = 0; i < b.numNodes(); ++i) {
        if (!b.isLeaf(i)) continue;
        ++leaves;
        const int depth = b.treeDepth(i, false);
        maxDepth = std::max(maxDepth, depth); minDepth = std::min(minDepth, depth); avgDepth += depth;
        const float avgRadiance = b.treeMean(i, false);
        maxAvgRadiance = std::max(maxAvgRadiance, avgRadiance); minAvgRadiance = std::min(minAvgRadiance, avgRadiance); avgAvgRadiance += avgRadiance;
        if (b.treeSize(i, false) > 1) {
            const size_t nodes = b.treeSize(i, false);
            maxNodes = std::max(maxNodes, nodes); minNodes = std::min(minNodes, nodes); avgNodes += nodes; ++nPointsNodes;
        }
        const float w = b.treeWeight(i, false);
        maxW = std::max(maxW, w); minW = std::min(minW, w); avgW += w;
        ++nPoints;
    }
    if (nPoints > 0) { avgDepth /= nPoints; avgAvgRadiance /= nPoints; if (nPointsNodes > 0) avgNodes /= nPointsNodes; avgW /= nPoints; }
    st.depth_min = minDepth; st.depth_max = maxDepth; st.depth_avg = avgDepth;
    st.mean_radiance_min = minAvgRadiance; st.mean_radiance_avg = avgAvgRadiance; st.mean_radiance_max = maxAvgRadiance;
    st.nodes_min = minNodes; st.nodes_max = maxNodes; st.nodes_avg = avgNodes;
    st.weight_min = minW; st.weight_avg = avgW; st.weight_max = maxW;
    st.s_tree_nodes = (uint32_t) b.numNodes(); st.s_tree_leaves = leaves;
}
}  // namespace ppgo
