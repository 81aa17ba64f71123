// sr_bvh.h -- deterministic host-side BVH2 builder (binned SAH, median fallback).
// Replaces the SpatialSubdivision constructor (Raytrace/SpatialSubdivision.cs:49-230,267-315) as
// the acceleration structure; unlike that kd-style tree every primitive lives in exactly one leaf
// and the depth is not capped at 15, so 1M-10M triangle scenes stay traceable.  The same input
// always yields the same node array and primitive order ("bit-identical layout across runs").
#pragma once
#include <cstdint>
#include <vector>

#include "sr_types.h"

namespace sr {

struct PrimBounds {   // FP32 bounds already rounded outward from the FP64 geometry
    float lo[3];
    float hi[3];
};

struct BvhBuild {
    std::vector<BvhNode> nodes;    // nodes[0] is the root and is always an internal node
    std::vector<int32_t> order;    // order[k] = input index of the k-th primitive in leaf order
    int32_t depth = 0;
    int32_t n_leaves = 0;
    float   root_lo[3] = {0, 0, 0}, root_hi[3] = {0, 0, 0};
};

float round_down(double x);
float round_up(double x);

// pad: absolute slack added to every node box (space units), see DESIGN.md "FP32 traversal".
// isect_cost: SAH cost of one primitive test relative to one node visit (two box tests + stack work).
void build_bvh(const std::vector<PrimBounds>& prims, float pad, int max_leaf, double isect_cost, BvhBuild* out);

}  // namespace sr
