// Model access through a loader (dense route, dense_kernel.cu; also used by the K1-split experiment, tools/experiments/k1_split).
#pragma once
#include "ode_models.cuh"

namespace magi {

// ---- model access through a loader (x(dd) = state component dd at this (time, chain)) so that Lorenz-96 with D = 64
// never materialises a 64-entry register array ----
template <int MODEL> struct DenseOde {
    static constexpr int K = Ode<MODEL>::K, KX = Ode<MODEL>::KX;
    static constexpr int SD = Ode<MODEL>::D;      // > 0: all components of a time point fit registers
    __device__ static void prepare(double* th) { Ode<MODEL>::prepare(th); }
    template <class X> __device__ static double f(int d, X x, const double* th, int) {
        double xa[Ode<MODEL>::D];
#pragma unroll
        for (int i = 0; i < Ode<MODEL>::D; ++i) xa[i] = x(i);
        return Ode<MODEL>::f(d, xa, th);
    }
    template <class X, class W> __device__ static void jx_col_sub(int j, X x, W w, const double* th, int, double& g) {
        double xa[Ode<MODEL>::D], wa[Ode<MODEL>::D];
#pragma unroll
        for (int i = 0; i < Ode<MODEL>::D; ++i) { xa[i] = x(i); wa[i] = w(i); }
        Ode<MODEL>::jx_col_sub(j, xa, th, wa, g);
    }
    template <class X> __device__ static void jth_row_sub(int p, X x, const double* th, int, double w, double* acc) {
        double xa[Ode<MODEL>::D];
#pragma unroll
        for (int i = 0; i < Ode<MODEL>::D; ++i) xa[i] = x(i);
        Ode<MODEL>::jth_row_sub(p, xa, th, w, acc);
    }
};
// Lorenz-96 (not in the reference; BASELINE config 4): x_i' = (x_{i+1} - x_{i-2}) x_{i-1} - x_i + F, cyclic.
template <> struct DenseOde<MAGI_MODEL_L96> {
    static constexpr int K = 1, KX = 1;
    static constexpr int SD = 0;
    __device__ static void prepare(double*) {}
    template <class X> __device__ static double f(int d, X x, const double* th, int D) {
        const int p1 = (d + 1) % D, m1 = (d + D - 1) % D, m2 = (d + D - 2) % D;
        return (x(p1) - x(m2)) * x(m1) - x(d) + th[0];
    }
    // column j of the Jacobian has four entries: rows j-1, j, j+1, j+2
    template <class X, class W> __device__ static void jx_col_sub(int j, X x, W w, const double*, int D, double& g) {
        const int jm2 = (j + D - 2) % D, jm1 = (j + D - 1) % D, jp1 = (j + 1) % D, jp2 = (j + 2) % D;
        g -= x(jm2) * w(jm1);                  // d f_{j-1} / d x_j = x_{j-2}
        g -= (-1.0) * w(j);                    // d f_j / d x_j = -1
        g -= (x(jp2) - x(jm1)) * w(jp1);       // d f_{j+1} / d x_j = x_{j+2} - x_{j-1}
        g -= (-x(jp1)) * w(jp2);               // d f_{j+2} / d x_j = -x_{j+1}
    }
    template <class X> __device__ static void jth_row_sub(int, X, const double*, int, double w, double* acc) { acc[0] -= w; }
};

}  // namespace magi
