// Multiresolution hash-grid encoding (Instant-NGP), F = 2 features per level, fp16 table.
//
// The reference delegates this to tinycudann (reference models/networks.py:36-56), which is NOT in
// /root/reference. The arithmetic restated here is tiny-cuda-nn's published GridEncoding
// ("[tcnn-memory]"; parity unpinned):
//   scale_l = exp2f(l*log2f(b))*N_min - 1 ; res_l = ceil(scale_l)+1 ; entries_l = min(align8(res^3), T)
//   pos = fmaf(scale_l, x, 0.5) ; g = floor(pos) ; w = pos - g
//   index(p) = dense ? p.x + p.y*res + p.z*res^2 : p.x ^ p.y*2654435761 ^ p.z*805459861 ;  index %= entries_l
//   value = sum over the 8 corners of prod(w or 1-w) * table[offset_l + index]   (fp32 accumulate)
#pragma once
#include "common.cuh"
#include "../../include/ngp_b200.h"

__device__ __forceinline__ uint32_t grid_corner_index(uint32_t px, uint32_t py, uint32_t pz, uint32_t res,
                                                      uint32_t entries, bool hashed) {
    uint32_t idx;
    if (hashed) {
        idx = px ^ (py * 2654435761u) ^ (pz * 805459861u);
        // hashed levels always have 2^log2_T entries
        idx &= (entries - 1u);
    } else {
        idx = px + py * res + pz * res * res;
        // For inputs in [0,1]^3 only the +1 corner on the upper boundary can run past the level and
        // then idx < 2*entries, so one conditional subtract IS the modulo. The clamp keeps inputs from
        // outside the unit cube memory-safe (tiny-cuda-nn would wrap them with a true %).
        idx = idx >= entries ? idx - entries : idx;
        idx = min(idx, entries - 1u);
    }
    return idx;
}

struct GridCell {
    uint32_t gx, gy, gz;
    float wx, wy, wz;
};

__device__ __forceinline__ GridCell grid_cell(float x01, float y01, float z01, float scale) {
    GridCell c;
    float p, f;
    p = fmaf(scale, x01, 0.5f); f = floorf(p); c.gx = (uint32_t)(int)f; c.wx = p - f;
    p = fmaf(scale, y01, 0.5f); f = floorf(p); c.gy = (uint32_t)(int)f; c.wy = p - f;
    p = fmaf(scale, z01, 0.5f); f = floorf(p); c.gz = (uint32_t)(int)f; c.wz = p - f;
    return c;
}

// Trilinear lookup of one (sample, level): 8 independent 4-byte gathers, then 8 FMAs per feature.
__device__ __forceinline__ float2 grid_lookup(const uint32_t* __restrict__ table /* half2 per entry */,
                                              const NgpGridMeta& m, int level, float x01, float y01, float z01) {
    const uint32_t res = m.res[level];
    const uint32_t off = m.offset[level];
    const uint32_t entries = m.offset[level + 1] - off;
    const bool hashed = (m.hashed_mask >> level) & 1u;
    const GridCell c = grid_cell(x01, y01, z01, m.scale[level]);
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t px = c.gx + (k & 1), py = c.gy + ((k >> 1) & 1), pz = c.gz + ((k >> 2) & 1);
        v[k] = __ldg(table + off + grid_corner_index(px, py, pz, res, entries, hashed));
    }
    float f0 = 0.f, f1 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float w = ((k & 1) ? c.wx : 1.0f - c.wx) * ((k & 2) ? c.wy : 1.0f - c.wy) * ((k & 4) ? c.wz : 1.0f - c.wz);
        const float2 t = unpack_half2(v[k]);
        f0 = fmaf(w, t.x, f0);
        f1 = fmaf(w, t.y, f1);
    }
    return make_float2(f0, f1);
}

// Scatter of one (sample, level) gradient into the fp32 gradient table: 8 vector reductions of 8 B.
__device__ __forceinline__ void grid_scatter(float* __restrict__ grad /* float2 per entry */, const NgpGridMeta& m, int level,
                                             float x01, float y01, float z01, float g0, float g1) {
    const uint32_t res = m.res[level];
    const uint32_t off = m.offset[level];
    const uint32_t entries = m.offset[level + 1] - off;
    const bool hashed = (m.hashed_mask >> level) & 1u;
    const GridCell c = grid_cell(x01, y01, z01, m.scale[level]);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t px = c.gx + (k & 1), py = c.gy + ((k >> 1) & 1), pz = c.gz + ((k >> 2) & 1);
        const uint32_t idx = off + grid_corner_index(px, py, pz, res, entries, hashed);
        const float w = ((k & 1) ? c.wx : 1.0f - c.wx) * ((k & 2) ? c.wy : 1.0f - c.wy) * ((k & 4) ? c.wz : 1.0f - c.wz);
        red_add_f32x2(grad + 2 * (size_t)idx, w * g0, w * g1);
    }
}

// Degree-4 real spherical harmonics of a unit vector (16 coefficients), tiny-cuda-nn's ordering and
// constants.
__device__ __forceinline__ void sh4(float x, float y, float z, float* __restrict__ o) {
    const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    o[0] = 0.28209479177387814f;
    o[1] = -0.48860251190291987f * y;
    o[2] = 0.48860251190291987f * z;
    o[3] = -0.48860251190291987f * x;
    o[4] = 1.0925484305920792f * xy;
    o[5] = -1.0925484305920792f * yz;
    o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    o[7] = -1.0925484305920792f * xz;
    o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
    o[10] = 2.8906114426405538f * xy * z;
    o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
    o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
    o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
    o[14] = 1.4453057213202769f * z * (x2 - y2);
    o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}
