// Occupancy-grid ray marcher core (device functions shared by every kernel that marches).
//
// Parity contract: the per-ray sequence of (t, dt, xyz) must be BIT-EXACT with the reference's
// raymarching_train_kernel / raymarching_test_kernel (reference models/csrc/raymarching.cu:166-280,
// :335-404). The reference is compiled with nvcc's default -fmad=true and no fast-math; which
// multiplies/adds get fused there is read off its SASS. Here every rounding
// step is spelled with an explicit intrinsic (__fmaf_rn / __fmul_rn / __fadd_rn / __fdiv_rn), so the
// result does not depend on how the compiler chooses to contract the surrounding code.
#pragma once
#include "common.cuh"

struct MarchConst {
    const uint8_t* __restrict__ bitfield;  // cascades * G^3 / 8 bytes, bit i%8 of byte i/8, i = mip*G^3 + morton(x,y,z)
    int cascades;
    int grid_size;
    int max_samples;
    uint32_t grid_size3;
    float scale;         // scene half extent (reference NGP.scale)
    float esf;           // exp_step_factor
    float dt_lo, dt_hi;  // clamp bounds of the step (reference raymarching.cu:11-13)
    float gs_f, gs_inv, gs_m1;
};

// dt_scale is what the reference hands to calc_dt as `scale`: NGP.scale for the train kernel
// (raymarching.cu:196,207,231) but `cascades` for the test kernel (raymarching.cu:370,399).
__device__ __forceinline__ MarchConst make_march_const(const uint8_t* bitfield, int cascades, int grid_size,
                                                       int max_samples, float scale, float esf, float dt_scale) {
    MarchConst c;
    c.bitfield = bitfield;
    c.cascades = cascades;
    c.grid_size = grid_size;
    c.max_samples = max_samples;
    c.grid_size3 = (uint32_t)grid_size * (uint32_t)grid_size * (uint32_t)grid_size;
    c.scale = scale;
    c.esf = esf;
    c.gs_f = (float)grid_size;
    c.gs_inv = __fdiv_rn(1.0f, c.gs_f);
    c.gs_m1 = __fadd_rn(c.gs_f, -1.0f);
    c.dt_lo = __fdiv_rn(1.73205080757f, (float)max_samples);
    c.dt_hi = __fdiv_rn(__fmul_rn(dt_scale, 3.46410161514f), c.gs_f);
    return c;
}

__device__ __forceinline__ float march_dt(float t, const MarchConst& c) {
    // clamp(t*esf, lo, hi) == fmaxf(lo, fminf(t*esf, hi))
    return fmaxf(c.dt_lo, fminf(__fmul_rn(t, c.esf), c.dt_hi));
}

// spread the low 10 bits of v so that there are two zero bits between consecutive bits
__device__ __host__ __forceinline__ uint32_t morton_spread10(uint32_t v) {
    v &= 0x000003ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__device__ __host__ __forceinline__ uint32_t morton_encode3(uint32_t x, uint32_t y, uint32_t z) {
    return morton_spread10(x) | (morton_spread10(y) << 1) | (morton_spread10(z) << 2);
}
__device__ __host__ __forceinline__ uint32_t morton_compact10(uint32_t v) {
    v &= 0x09249249u;
    v = (v | (v >> 2)) & 0x030c30c3u;
    v = (v | (v >> 4)) & 0x0300f00fu;
    v = (v | (v >> 8)) & 0x030000ffu;
    v = (v | (v >> 16)) & 0x000003ffu;
    return v;
}

struct MarchRay {
    float ox, oy, oz;
    float dx, dy, dz;
    float ix, iy, iz;  // IEEE 1/d
    float sx, sy, sz;  // copysign(1, d)
};

__device__ __forceinline__ MarchRay make_march_ray(float ox, float oy, float oz, float dx, float dy, float dz) {
    MarchRay r;
    r.ox = ox; r.oy = oy; r.oz = oz;
    r.dx = dx; r.dy = dy; r.dz = dz;
    r.ix = __fdiv_rn(1.0f, dx); r.iy = __fdiv_rn(1.0f, dy); r.iz = __fdiv_rn(1.0f, dz);
    r.sx = copysignf(1.0f, dx); r.sy = copysignf(1.0f, dy); r.sz = copysignf(1.0f, dz);
    return r;
}

// Slab test against one axis-aligned box (reference intersection.cu:5-22, :45-52), followed by the
// near-plane clamp render() applies (reference rendering.py:29). Returns (t1,t2), (-1,-1) on a miss.
__device__ __forceinline__ float2 ray_aabb(const MarchRay& r, float cx, float cy, float cz, float hx, float hy, float hz) {
    const float ax = __fmul_rn(__fsub_rn(__fsub_rn(cx, hx), r.ox), r.ix);
    const float bx = __fmul_rn(__fsub_rn(__fadd_rn(cx, hx), r.ox), r.ix);
    const float ay = __fmul_rn(__fsub_rn(__fsub_rn(cy, hy), r.oy), r.iy);
    const float by = __fmul_rn(__fsub_rn(__fadd_rn(cy, hy), r.oy), r.iy);
    const float az = __fmul_rn(__fsub_rn(__fsub_rn(cz, hz), r.oz), r.iz);
    const float bz = __fmul_rn(__fsub_rn(__fadd_rn(cz, hz), r.oz), r.iz);
    const float t1 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    const float t2 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    if (t1 > t2) return make_float2(-1.0f, -1.0f);
    return make_float2(t1, t2);
}

// One visit of the marcher at parameter t. Returns true when the cell under the ray is occupied
// (then (x,y,z,dt) describe the sample and the caller advances t += dt); otherwise t has already
// been advanced past the empty cell.
__device__ __forceinline__ bool march_visit(const MarchRay& r, const MarchConst& c, float& t,
                                            float& x, float& y, float& z, float& dt) {
    x = __fmaf_rn(r.dx, t, r.ox);
    y = __fmaf_rn(r.dy, t, r.oy);
    z = __fmaf_rn(r.dz, t, r.oz);
    dt = march_dt(t, c);

    int e_pos, e_dt;
    frexpf(fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z))), &e_pos);
    frexpf(__fmul_rn(dt, c.gs_f), &e_dt);
    const int mip_pos = min(c.cascades - 1, max(0, e_pos + 1));
    const int mip_dt = min(c.cascades - 1, max(0, e_dt));
    const int mip = max(mip_pos, mip_dt);

    const float mip_bound = fminf(scalbnf(1.0f, mip - 1), c.scale);
    const float mip_bound_inv = __fdiv_rn(1.0f, mip_bound);

    float vx = __fmul_rn(__fmul_rn(__fmaf_rn(x, mip_bound_inv, 1.0f), 0.5f), c.gs_f);
    float vy = __fmul_rn(__fmul_rn(__fmaf_rn(y, mip_bound_inv, 1.0f), 0.5f), c.gs_f);
    float vz = __fmul_rn(__fmul_rn(__fmaf_rn(z, mip_bound_inv, 1.0f), 0.5f), c.gs_f);
    const int nx = (int)fmaxf(0.0f, fminf(vx, c.gs_m1));
    const int ny = (int)fmaxf(0.0f, fminf(vy, c.gs_m1));
    const int nz = (int)fmaxf(0.0f, fminf(vz, c.gs_m1));

    const uint32_t idx = (uint32_t)mip * c.grid_size3 + morton_encode3((uint32_t)nx, (uint32_t)ny, (uint32_t)nz);
    const bool occ = (__ldg(c.bitfield + (idx >> 3)) >> (idx & 7u)) & 1u;
    if (occ) return true;

    // distance to the exit face of this cell along each axis, then step-quantised advance
    float a;
    a = __fmaf_rn(r.sx, 0.5f, __fadd_rn((float)nx, 0.5f));
    a = __fmaf_rn(__fmul_rn(a, c.gs_inv), 2.0f, -1.0f);
    const float tx = __fmul_rn(__fmaf_rn(mip_bound, a, -x), r.ix);
    a = __fmaf_rn(r.sy, 0.5f, __fadd_rn((float)ny, 0.5f));
    a = __fmaf_rn(__fmul_rn(a, c.gs_inv), 2.0f, -1.0f);
    const float ty = __fmul_rn(__fmaf_rn(mip_bound, a, -y), r.iy);
    a = __fmaf_rn(r.sz, 0.5f, __fadd_rn((float)nz, 0.5f));
    a = __fmaf_rn(__fmul_rn(a, c.gs_inv), 2.0f, -1.0f);
    const float tz = __fmul_rn(__fmaf_rn(mip_bound, a, -z), r.iz);

    const float t_target = __fadd_rn(t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
    do {
        t = __fadd_rn(t, march_dt(t, c));
    } while (t < t_target);
    return false;
}

// Train-time start jitter (reference raymarching.cu:195-198): only for rays that hit the box.
__device__ __forceinline__ float march_jitter(float t1, float noise, const MarchConst& c) {
    if (t1 >= 0.0f) t1 = __fmaf_rn(march_dt(t1, c), noise, t1);
    return t1;
}
