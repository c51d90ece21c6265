// Front-to-back alpha compositing along one ray by one warp (32 samples per trip, transmittance as
// a warp-wide inclusive product scan with a carry between trips).
//
// Semantics follow reference models/csrc/volumerendering.cu:6-45 (forward) and :87-151 (backward):
//   a_i = 1 - __expf(-sigma_i*delta_i);  w_i = a_i * T_{i-1};  T_i = T_{i-1}*(1-a_i)
//   the sample on which T first drops to <= T_threshold IS accumulated, later ones are not,
//   total_samples excludes that terminating sample, ws stays 0 past it.
// The reference accumulates serially per ray; a warp scan re-associates the fp32 sums/products
// (~1e-7 relative), which is inside the 1e-4 relative budget of the parity tests but not bit-exact.
#pragma once
#include "common.cuh"

struct CompositeOut {
    float r, g, b, depth, opacity;
    int n_composited;   // samples that received a non-zero weight slot (includes the terminating one)
    int total_samples;  // reference `total_samples[ray]`
};

// Sample accessors are functors so the same code serves the ragged (train) layout and fused kernels.
//   sig(i), dlt(i), tt(i) -> float ; col(i) -> float3 ; put_w(i, w) stores ws (may be a no-op)
template <class FSig, class FDlt, class FT, class FCol, class FPutW>
__device__ __forceinline__ CompositeOut composite_ray_warp(int n, float T_threshold, int lane,
                                                           FSig sig, FDlt dlt, FT tt, FCol col, FPutW put_w) {
    float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_d = 0.f, acc_o = 0.f;
    float T_carry = 1.0f;
    int n_comp = 0;
    bool done = false;
    int base = 0;
    for (; base < n && !done; base += 32) {
        const int i = base + lane;
        const bool valid = i < n;
        float a = 0.f;
        if (valid) a = 1.0f - __expf(-(sig(i) * dlt(i)));
        const float om = 1.0f - a;
        const float T_inc = warp_scan_mul(om, lane) * T_carry;
        float T_exc = __shfl_up_sync(0xffffffffu, T_inc, 1);
        if (lane == 0) T_exc = T_carry;
        // composited iff the transmittance BEFORE this sample is still above the threshold
        const bool comp = valid && (i == 0 || T_exc > T_threshold);
        const float w = comp ? a * T_exc : 0.f;
        if (comp) {
            const float3 c = col(i);
            acc_r = fmaf(w, c.x, acc_r);
            acc_g = fmaf(w, c.y, acc_g);
            acc_b = fmaf(w, c.z, acc_b);
            acc_d = fmaf(w, tt(i), acc_d);
            acc_o += w;
        }
        if (valid) put_w(i, w);
        const unsigned m = __ballot_sync(0xffffffffu, comp);
        n_comp += __popc(m);
        // terminated inside this trip?
        const unsigned term = __ballot_sync(0xffffffffu, valid && !(T_inc > T_threshold));
        done = term != 0u;
        T_carry = __shfl_sync(0xffffffffu, T_inc, 31);
    }
    // zero the weights of samples never visited (past the terminating trip)
    for (int i = base + lane; i < n; i += 32) put_w(i, 0.f);

    CompositeOut o;
    o.r = warp_sum(acc_r);
    o.g = warp_sum(acc_g);
    o.b = warp_sum(acc_b);
    o.depth = warp_sum(acc_d);
    o.opacity = warp_sum(acc_o);
    o.n_composited = n_comp;
    o.total_samples = done ? n_comp - 1 : n_comp;
    return o;
}

// Backward of the above for one ray by one warp. Returns the number of composited samples (the leading samples of
// the ray, terminating one included): only those can receive a non-zero gradient.
//   dsig(i, v) / dcol(i, float3) store the per-sample gradients (0 for samples past termination).
//   dws(i) is dL/dws_i (pass a functor returning 0 when no loss touches ws); wsv(i) = forward ws_i.
template <class FSig, class FDlt, class FT, class FCol, class FDws, class FWs, class FPutS, class FPutC>
__device__ __forceinline__ int composite_ray_warp_bwd(int n, float T_threshold, int lane,
                                                       float dO, float dD, float3 dC,
                                                       float O, float D, float3 C,
                                                       FSig sig, FDlt dlt, FT tt, FCol col, FDws dws, FWs wsv,
                                                       FPutS dsig, FPutC dcol) {
    // total of dL/dws * ws over the ray (ws is 0 past termination)
    float s_tot = 0.f;
    for (int i = lane; i < n; i += 32) s_tot = fmaf(dws(i), wsv(i), s_tot);
    s_tot = warp_sum(s_tot);

    float T_carry = 1.0f;
    float pr = 0.f, pg = 0.f, pb = 0.f, pd = 0.f, ps = 0.f;  // inclusive prefixes carried between trips
    bool done = false;
    int base = 0;
    int n_comp = 0;
    for (; base < n && !done; base += 32) {
        const int i = base + lane;
        const bool valid = i < n;
        float a = 0.f, de = 0.f, ti = 0.f, dw = 0.f, wv = 0.f;
        float3 c = make_float3(0.f, 0.f, 0.f);
        if (valid) {
            de = dlt(i);
            a = 1.0f - __expf(-(sig(i) * de));
            c = col(i);
            ti = tt(i);
            dw = dws(i);
            wv = wsv(i);
        }
        const float om = 1.0f - a;
        const float T_inc = warp_scan_mul(om, lane) * T_carry;
        float T_exc = __shfl_up_sync(0xffffffffu, T_inc, 1);
        if (lane == 0) T_exc = T_carry;
        const bool comp = valid && (i == 0 || T_exc > T_threshold);
        const float w = comp ? a * T_exc : 0.f;
        const float r_inc = warp_scan_add(w * c.x, lane) + pr;
        const float g_inc = warp_scan_add(w * c.y, lane) + pg;
        const float b_inc = warp_scan_add(w * c.z, lane) + pb;
        const float d_inc = warp_scan_add(w * ti, lane) + pd;
        const float s_inc = warp_scan_add(dw * wv, lane) + ps;
        if (valid) {
            if (comp) {
                dcol(i, make_float3(dC.x * w, dC.y * w, dC.z * w));
                const float g = dC.x * (c.x * T_inc - (C.x - r_inc)) +
                                dC.y * (c.y * T_inc - (C.y - g_inc)) +
                                dC.z * (c.z * T_inc - (C.z - b_inc)) +
                                dO * (1.0f - O) +
                                dD * (ti * T_inc - (D - d_inc)) +
                                (T_inc * dw - (s_tot - s_inc));
                dsig(i, de * g);
            } else {
                dcol(i, make_float3(0.f, 0.f, 0.f));
                dsig(i, 0.f);
            }
        }
        n_comp += __popc(__ballot_sync(0xffffffffu, comp));
        const unsigned term = __ballot_sync(0xffffffffu, valid && !(T_inc > T_threshold));
        done = term != 0u;
        T_carry = __shfl_sync(0xffffffffu, T_inc, 31);
        pr = __shfl_sync(0xffffffffu, r_inc, 31);
        pg = __shfl_sync(0xffffffffu, g_inc, 31);
        pb = __shfl_sync(0xffffffffu, b_inc, 31);
        pd = __shfl_sync(0xffffffffu, d_inc, 31);
        ps = __shfl_sync(0xffffffffu, s_inc, 31);
    }
    for (int i = base + lane; i < n; i += 32) {
        dcol(i, make_float3(0.f, 0.f, 0.f));
        dsig(i, 0.f);
    }
    return n_comp;
}
