// Device functors for the compiled ODE model registry (reference: src/ode_models.jl).
// Each model exposes, for a state x[D] and parameters th[K] at one time point:
//   f(d, x, th)                      component d of the right-hand side
//   jx_col_sub(j, x, th, w, g)       g -= sum_p (df_p/dx_j) * w[p], p ascending -- the order the reference's
//                                    dimension loop accumulates (src/likelihoods.jl:214-216)
//   jth_row_sub(p, x, th, w, acc)    acc[q] -= (df_p/dth_q) * w for q < K        (src/likelihoods.jl:219-221)
// x holds ALL D components (small-D models).  th holds the K parameters followed by KX - K per-chain invariants filled by
// prepare(th) ONCE per chain (an FP64 division is a ~30-instruction dependent chain on the FP64 pipe the DMMAs need; under
// register pressure the compiler re-evaluates 1/c at every time point instead of hoisting it);
// the only deliberate deviation from the reference's arithmetic is V^3/3.0 written as a multiplication by 1/3
// (<= 1 ulp), because an FP64 division costs ~10 DFMA-pipe slots per time point.
#pragma once
#include "../../include/magi_b200.h"

namespace magi {

template <int MODEL> struct Ode;

// ---- FitzHugh-Nagumo: src/ode_models.jl:39-47, 248-262, 274-299 ----
template <> struct Ode<MAGI_MODEL_FN> {
    static constexpr int D = 2, K = 3, KX = 7;      // th[3] = 1/c, th[4] = -1/c, th[5] = -b/c, th[6] = 1/(c*c)
    __device__ __forceinline__ static void prepare(double* th) {
        const double b = th[1], c = th[2];
        th[3] = 1.0 / c; th[4] = -1.0 / c; th[5] = -b / c; th[6] = 1.0 / (c * c);
    }
    __device__ __forceinline__ static double f(int d, const double* x, const double* th) {
        const double V = x[0], R = x[1], a = th[0], b = th[1], c = th[2];
        if (d == 0) return c * (V - (V * V * V) * (1.0 / 3.0) + R);
        return th[4] * (V - a + b * R);                                   // (-1.0/c) * (V - a + b R)
    }
    __device__ __forceinline__ static void jx_col_sub(int j, const double* x, const double* th, const double* w, double& g) {
        const double V = x[0], c = th[2];
        if (j == 0) { g -= (c * (1.0 - V * V)) * w[0]; g -= th[4] * w[1]; }
        else        { g -= c * w[0];                   g -= th[5] * w[1]; }
    }
    __device__ __forceinline__ static void jth_row_sub(int p, const double* x, const double* th, double w, double* acc) {
        const double V = x[0], R = x[1], a = th[0], b = th[1];
        if (p == 0) { acc[2] -= (V - (V * V * V) * (1.0 / 3.0) + R) * w; }
        else {
            acc[0] -= th[3] * w;
            acc[1] -= (-R * th[3]) * w;                                   // -R/c (<= 1 ulp from the reference's division)
            acc[2] -= (th[6] * (V - a + b * R)) * w;
        }
    }
};

// ---- Hes1: src/ode_models.jl:60-70, 312-336, 349-378 ----
template <> struct Ode<MAGI_MODEL_HES1> {
    static constexpr int D = 3, K = 7, KX = 7;
    __device__ __forceinline__ static void prepare(double*) {}
    __device__ __forceinline__ static double f(int d, const double* x, const double* p) {
        const double P = x[0], M = x[1], H = x[2];
        if (d == 0) return -p[0] * P * H + p[1] * M - p[2] * P;
        if (d == 1) return -p[3] * M + p[4] / (1.0 + P * P);
        return -p[0] * P * H + p[5] / (1.0 + P * P) - p[6] * H;
    }
    __device__ __forceinline__ static void jx_col_sub(int j, const double* x, const double* p, const double* w, double& g) {
        const double P = x[0], H = x[2];
        if (j == 0) {
            const double opp = 1.0 + P * P, r = (2.0 * P) / (opp * opp);
            g -= (-p[0] * H - p[2]) * w[0];
            g -= (-p[4] * r) * w[1];
            g -= (-p[0] * H - p[5] * r) * w[2];
        } else if (j == 1) {
            g -= p[1] * w[0]; g -= (-p[3]) * w[1]; g -= 0.0 * w[2];
        } else {
            g -= (-p[0] * P) * w[0]; g -= 0.0 * w[1]; g -= (-p[0] * P - p[6]) * w[2];
        }
    }
    __device__ __forceinline__ static void jth_row_sub(int r, const double* x, const double* p, double w, double* acc) {
        const double P = x[0], M = x[1], H = x[2];
        if (r == 0) { acc[0] -= (-P * H) * w; acc[1] -= M * w; acc[2] -= (-P) * w; }
        else if (r == 1) { acc[3] -= (-M) * w; acc[4] -= (1.0 / (1.0 + P * P)) * w; }
        else { acc[0] -= (-P * H) * w; acc[5] -= (1.0 / (1.0 + P * P)) * w; acc[6] -= (-H) * w; }
    }
};

// ---- Lotka-Volterra (not in the reference; BASELINE config 3): x' = a x - b x y, y' = d x y - g y ----
template <> struct Ode<MAGI_MODEL_LV> {
    static constexpr int D = 2, K = 4, KX = 4;
    __device__ __forceinline__ static void prepare(double*) {}
    __device__ __forceinline__ static double f(int d, const double* x, const double* th) {
        if (d == 0) return th[0] * x[0] - th[1] * x[0] * x[1];
        return th[2] * x[0] * x[1] - th[3] * x[1];
    }
    __device__ __forceinline__ static void jx_col_sub(int j, const double* x, const double* th, const double* w, double& g) {
        if (j == 0) { g -= (th[0] - th[1] * x[1]) * w[0]; g -= (th[2] * x[1]) * w[1]; }
        else        { g -= (-th[1] * x[0]) * w[0];        g -= (th[2] * x[0] - th[3]) * w[1]; }
    }
    __device__ __forceinline__ static void jth_row_sub(int p, const double* x, const double* th, double w, double* acc) {
        if (p == 0) { acc[0] -= x[0] * w; acc[1] -= (-x[0] * x[1]) * w; }
        else        { acc[2] -= (x[0] * x[1]) * w; acc[3] -= (-x[1]) * w; }
    }
};

// ---- Hes1 in log coordinates: src/ode_models.jl:83-103 (+ fixed-parameter variants :116-165).
// The reference ships no Jacobians for these; the ones below are derived and FD-checked in tests. ----
template <int VARIANT> struct OdeHes1Log {   // 0: 7 params; 1: gamma fixed 0.3 (6 params); 2: f fixed 20 (6 params)
    static constexpr int D = 3, K = (VARIANT == 0 ? 7 : 6), KX = K;
    __device__ __forceinline__ static void prepare(double*) {}
    __device__ __forceinline__ static void unpack(const double* p, double* q) {
        q[0] = p[0]; q[1] = p[1]; q[2] = p[2]; q[3] = p[3]; q[4] = p[4];
        if (VARIANT == 0) { q[5] = p[5]; q[6] = p[6]; }
        else if (VARIANT == 1) { q[5] = p[5]; q[6] = 0.3; }
        else { q[5] = 20.0; q[6] = p[5]; }
    }
    __device__ __forceinline__ static double f(int d, const double* x, const double* p) {
        double q[7]; unpack(p, q);
        const double P = exp(x[0]), M = exp(x[1]), H = exp(x[2]);
        const double opp = 1.0 + P * P;
        if (d == 0) return -q[0] * H + q[1] * M / P - q[2];
        if (d == 1) return -q[3] + q[4] / (opp * M);
        return -q[0] * P + q[5] / (opp * H) - q[6];
    }
    // derivatives w.r.t. the log states
    __device__ __forceinline__ static void jx_col_sub(int j, const double* x, const double* p, const double* w, double& g) {
        double q[7]; unpack(p, q);
        const double P = exp(x[0]), M = exp(x[1]), H = exp(x[2]);
        const double opp = 1.0 + P * P;
        if (j == 0) {
            g -= (-q[1] * M / P) * w[0];
            g -= (-q[4] * 2.0 * P * P / (opp * opp * M)) * w[1];
            g -= (-q[0] * P - q[5] * 2.0 * P * P / (opp * opp * H)) * w[2];
        } else if (j == 1) {
            g -= (q[1] * M / P) * w[0];
            g -= (-q[4] / (opp * M)) * w[1];
        } else {
            g -= (-q[0] * H) * w[0];
            g -= (-q[5] / (opp * H)) * w[2];
        }
    }
    __device__ __forceinline__ static void jth_row_sub(int r, const double* x, const double* p, double w, double* acc) {
        const double P = exp(x[0]), M = exp(x[1]), H = exp(x[2]);
        const double opp = 1.0 + P * P;
        if (r == 0) { acc[0] -= (-H) * w; acc[1] -= (M / P) * w; acc[2] -= (-1.0) * w; }
        else if (r == 1) { acc[3] -= (-1.0) * w; acc[4] -= (1.0 / (opp * M)) * w; }
        else {
            acc[0] -= (-P) * w;
            if (VARIANT == 0) { acc[5] -= (1.0 / (opp * H)) * w; acc[6] -= (-1.0) * w; }
            else if (VARIANT == 1) { acc[5] -= (1.0 / (opp * H)) * w; }
            else { acc[5] -= (-1.0) * w; }
        }
    }
};
template <> struct Ode<MAGI_MODEL_HES1LOG> : OdeHes1Log<0> {};
template <> struct Ode<MAGI_MODEL_HES1LOG_FIXG> : OdeHes1Log<1> {};
template <> struct Ode<MAGI_MODEL_HES1LOG_FIXF> : OdeHes1Log<2> {};

// ---- HIV in log coordinates: src/ode_models.jl:178-207 (Jacobians derived) ----
template <> struct Ode<MAGI_MODEL_HIV> {
    static constexpr int D = 4, K = 9, KX = 9;
    __device__ __forceinline__ static void prepare(double*) {}
    __device__ __forceinline__ static double f(int d, const double* x, const double* p) {
        const double T = exp(x[0]), Tm = exp(x[1]), Tw = exp(x[2]), Tmw = exp(x[3]);
        const double sf = 1e-6;
        if (d == 0) return p[0] - sf * p[1] * Tm - sf * p[2] * Tw - sf * p[3] * Tmw;
        if (d == 1) return p[6] + sf * p[1] * T - sf * p[4] * Tw + sf * 0.25 * p[3] * Tmw * T / Tm;
        if (d == 2) return p[7] + sf * p[2] * T - sf * p[5] * Tm + sf * 0.25 * p[3] * Tmw * T / Tw;
        return p[8] + 0.5 * sf * p[3] * T + (sf * p[4] + sf * p[5]) * Tw * Tm / Tmw;
    }
    __device__ __forceinline__ static void jx_col_sub(int j, const double* x, const double* p, const double* w, double& g) {
        const double T = exp(x[0]), Tm = exp(x[1]), Tw = exp(x[2]), Tmw = exp(x[3]);
        const double sf = 1e-6;
        const double a1 = sf * 0.25 * p[3] * Tmw * T / Tm, a2 = sf * 0.25 * p[3] * Tmw * T / Tw;
        const double a3 = (sf * p[4] + sf * p[5]) * Tw * Tm / Tmw;
        if (j == 0) {        // d/d logT
            g -= 0.0 * w[0]; g -= (sf * p[1] * T + a1) * w[1]; g -= (sf * p[2] * T + a2) * w[2]; g -= (0.5 * sf * p[3] * T) * w[3];
        } else if (j == 1) { // d/d logTm
            g -= (-sf * p[1] * Tm) * w[0]; g -= (-a1) * w[1]; g -= (-sf * p[5] * Tm) * w[2]; g -= a3 * w[3];
        } else if (j == 2) { // d/d logTw
            g -= (-sf * p[2] * Tw) * w[0]; g -= (-sf * p[4] * Tw) * w[1]; g -= (-a2) * w[2]; g -= a3 * w[3];
        } else {             // d/d logTmw
            g -= (-sf * p[3] * Tmw) * w[0]; g -= a1 * w[1]; g -= a2 * w[2]; g -= (-a3) * w[3];
        }
    }
    __device__ __forceinline__ static void jth_row_sub(int r, const double* x, const double* p, double w, double* acc) {
        const double T = exp(x[0]), Tm = exp(x[1]), Tw = exp(x[2]), Tmw = exp(x[3]);
        const double sf = 1e-6;
        if (r == 0) { acc[0] -= w; acc[1] -= (-sf * Tm) * w; acc[2] -= (-sf * Tw) * w; acc[3] -= (-sf * Tmw) * w; }
        else if (r == 1) { acc[6] -= w; acc[1] -= (sf * T) * w; acc[4] -= (-sf * Tw) * w; acc[3] -= (sf * 0.25 * Tmw * T / Tm) * w; }
        else if (r == 2) { acc[7] -= w; acc[2] -= (sf * T) * w; acc[5] -= (-sf * Tm) * w; acc[3] -= (sf * 0.25 * Tmw * T / Tw) * w; }
        else { acc[8] -= w; acc[3] -= (0.5 * sf * T) * w; acc[4] -= (sf * Tw * Tm / Tmw) * w; acc[5] -= (sf * Tw * Tm / Tmw) * w; }
    }
};

// ---- protein transduction: src/ode_models.jl:219-233 (Jacobians derived) ----
template <> struct Ode<MAGI_MODEL_PTRANS> {
    static constexpr int D = 5, K = 6, KX = 6;
    __device__ __forceinline__ static void prepare(double*) {}
    __device__ __forceinline__ static double f(int d, const double* x, const double* p) {
        const double S = x[0], R = x[2], RS = x[3], RPP = x[4];
        if (d == 0) return -p[0] * S - p[1] * S * R + p[2] * RS;
        if (d == 1) return p[0] * S;
        if (d == 2) return -p[1] * S * R + p[2] * RS + p[4] * RPP / (p[5] + RPP);
        if (d == 3) return p[1] * S * R - p[2] * RS - p[3] * RS;
        return p[3] * RS - p[4] * RPP / (p[5] + RPP);
    }
    __device__ __forceinline__ static void jx_col_sub(int j, const double* x, const double* p, const double* w, double& g) {
        const double S = x[0], R = x[2], RPP = x[4];
        const double mm = p[4] * p[5] / ((p[5] + RPP) * (p[5] + RPP));   // d/dRPP of p5*RPP/(p6+RPP)
        if (j == 0) { g -= (-p[0] - p[1] * R) * w[0]; g -= p[0] * w[1]; g -= (-p[1] * R) * w[2]; g -= (p[1] * R) * w[3]; g -= 0.0 * w[4]; }
        else if (j == 1) { g -= 0.0 * w[0]; }
        else if (j == 2) { g -= (-p[1] * S) * w[0]; g -= 0.0 * w[1]; g -= (-p[1] * S) * w[2]; g -= (p[1] * S) * w[3]; }
        else if (j == 3) { g -= p[2] * w[0]; g -= 0.0 * w[1]; g -= p[2] * w[2]; g -= (-p[2] - p[3]) * w[3]; g -= p[3] * w[4]; }
        else { g -= mm * w[2]; g -= (-mm) * w[4]; }
    }
    __device__ __forceinline__ static void jth_row_sub(int r, const double* x, const double* p, double w, double* acc) {
        const double S = x[0], R = x[2], RS = x[3], RPP = x[4];
        const double den = p[5] + RPP;
        if (r == 0) { acc[0] -= (-S) * w; acc[1] -= (-S * R) * w; acc[2] -= RS * w; }
        else if (r == 1) { acc[0] -= S * w; }
        else if (r == 2) { acc[1] -= (-S * R) * w; acc[2] -= RS * w; acc[4] -= (RPP / den) * w; acc[5] -= (-p[4] * RPP / (den * den)) * w; }
        else if (r == 3) { acc[1] -= (S * R) * w; acc[2] -= (-RS) * w; acc[3] -= (-RS) * w; }
        else { acc[3] -= RS * w; acc[4] -= (-RPP / den) * w; acc[5] -= (p[4] * RPP / (den * den)) * w; }
    }
};

}  // namespace magi
