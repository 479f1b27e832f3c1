// pn_taylor.cuh -- scalar helpers shared by traced right-hand sides, and truncated Taylor-series
// arithmetic used for the on-device TaylorModeInit.
//
// Reference behaviour replaced: src/initialization/autodiffinit.jl:52-67 (get_derivatives with
// TaylorIntegration.jetcoeffs!) -- the exact derivatives u^(k)(t0), k = 0..q, that become the initial
// filter mean (after conditioning on them Sigma_0.R == 0 exactly).
//
// NVRTC-compatible: no standard headers.
#ifndef PN_TAYLOR_CUH
#define PN_TAYLOR_CUH

#ifndef PN_HD
#define PN_HD __device__ __forceinline__
#endif

// ---- double overloads used by generated code (tracer.py::_DevicePrinter) ----
PN_HD double pn_sq(double x) { return x * x; }
PN_HD double pn_ipow(double x, int n) {
    double r = 1.0;
    for (int i = 0; i < n; i++) r *= x;
    return r;
}
PN_HD double pn_sqrt(double x) { return sqrt(x); }
PN_HD double pn_pow(double x, double r) { return pow(x, r); }
// x^(k/2), k odd: sqrt(x)^k (k > 0) or (1/sqrt(x))^|k| -- replaces pow() in traced right-hand sides (e.g. r^-3)
PN_HD double pn_pow_half(double x, int k) {
#ifdef __CUDA_ARCH__
    const double r = (k > 0) ? sqrt(x) : rsqrt(x);
#else
    const double r = (k > 0) ? sqrt(x) : 1.0 / sqrt(x);  // host build of the header (tests/test_host.py)
#endif
    const int m = (k > 0) ? k : -k;
    double y = r;
    for (int i = 1; i < m; i++) y *= r;
    return y;
}
PN_HD double pn_exp(double x) { return exp(x); }
PN_HD double pn_log(double x) { return log(x); }
PN_HD double pn_sin(double x) { return sin(x); }
PN_HD double pn_cos(double x) { return cos(x); }
PN_HD double pn_tanh(double x) { return tanh(x); }

// ---- truncated Taylor polynomial of degree N (coefficients, not derivatives) ----
template <int N>
struct PnTaylor {
    double c[N + 1];
    PN_HD PnTaylor() {
#pragma unroll
        for (int k = 0; k <= N; k++) c[k] = 0.0;
    }
    PN_HD PnTaylor(double x) {
        c[0] = x;
#pragma unroll
        for (int k = 1; k <= N; k++) c[k] = 0.0;
    }
};

template <int N>
PN_HD PnTaylor<N> operator+(const PnTaylor<N>& a, const PnTaylor<N>& b) {
    PnTaylor<N> r;
#pragma unroll
    for (int k = 0; k <= N; k++) r.c[k] = a.c[k] + b.c[k];
    return r;
}
template <int N>
PN_HD PnTaylor<N> operator-(const PnTaylor<N>& a, const PnTaylor<N>& b) {
    PnTaylor<N> r;
#pragma unroll
    for (int k = 0; k <= N; k++) r.c[k] = a.c[k] - b.c[k];
    return r;
}
template <int N>
PN_HD PnTaylor<N> operator-(const PnTaylor<N>& a) {
    PnTaylor<N> r;
#pragma unroll
    for (int k = 0; k <= N; k++) r.c[k] = -a.c[k];
    return r;
}
template <int N>
PN_HD PnTaylor<N> operator+(const PnTaylor<N>& a, double b) {
    PnTaylor<N> r = a;
    r.c[0] += b;
    return r;
}
template <int N>
PN_HD PnTaylor<N> operator+(double b, const PnTaylor<N>& a) { return a + b; }
template <int N>
PN_HD PnTaylor<N> operator-(const PnTaylor<N>& a, double b) {
    PnTaylor<N> r = a;
    r.c[0] -= b;
    return r;
}
template <int N>
PN_HD PnTaylor<N> operator-(double b, const PnTaylor<N>& a) {
    PnTaylor<N> r = -a;
    r.c[0] += b;
    return r;
}
template <int N>
PN_HD PnTaylor<N> operator*(const PnTaylor<N>& a, double b) {
    PnTaylor<N> r;
#pragma unroll
    for (int k = 0; k <= N; k++) r.c[k] = a.c[k] * b;
    return r;
}
template <int N>
PN_HD PnTaylor<N> operator*(double b, const PnTaylor<N>& a) { return a * b; }
template <int N>
PN_HD PnTaylor<N> operator/(const PnTaylor<N>& a, double b) {
    PnTaylor<N> r;
#pragma unroll
    for (int k = 0; k <= N; k++) r.c[k] = a.c[k] / b;
    return r;
}
template <int N>
PN_HD PnTaylor<N> operator*(const PnTaylor<N>& a, const PnTaylor<N>& b) {
    PnTaylor<N> r;
#pragma unroll
    for (int k = 0; k <= N; k++) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j <= N; j++)
            if (j <= k) s += a.c[j] * b.c[j <= k ? k - j : 0];
        r.c[k] = s;
    }
    return r;
}
template <int N>
PN_HD PnTaylor<N> operator/(const PnTaylor<N>& a, const PnTaylor<N>& b) {
    PnTaylor<N> r;
#pragma unroll
    for (int k = 0; k <= N; k++) {
        double s = a.c[k];
#pragma unroll
        for (int j = 0; j <= N; j++)
            if (j < k) s -= r.c[j] * b.c[j < k ? k - j : 0];
        r.c[k] = s / b.c[0];
    }
    return r;
}
template <int N>
PN_HD PnTaylor<N> operator/(double a, const PnTaylor<N>& b) { return PnTaylor<N>(a) / b; }

template <int N>
PN_HD PnTaylor<N> pn_sq(const PnTaylor<N>& a) { return a * a; }
template <int N>
PN_HD PnTaylor<N> pn_ipow(const PnTaylor<N>& a, int n) {
    PnTaylor<N> r(1.0);
    for (int i = 0; i < n; i++) r = r * a;
    return r;
}
template <int N>
PN_HD PnTaylor<N> pn_sqrt(const PnTaylor<N>& a) {
    PnTaylor<N> r;
    r.c[0] = sqrt(a.c[0]);
#pragma unroll
    for (int k = 1; k <= N; k++) {
        double s = a.c[k];
#pragma unroll
        for (int j = 1; j <= N; j++)
            if (j < k) s -= r.c[j] * r.c[j < k ? k - j : 0];
        r.c[k] = s / (2.0 * r.c[0]);
    }
    return r;
}
// a^r for real r, a_0 != 0:  k a_0 p_k = sum_{j=1..k} (r j - (k - j)) a_j p_{k-j}
template <int N>
PN_HD PnTaylor<N> pn_pow(const PnTaylor<N>& a, double r) {
    PnTaylor<N> p;
    p.c[0] = pow(a.c[0], r);
#pragma unroll
    for (int k = 1; k <= N; k++) {
        double s = 0.0;
#pragma unroll
        for (int j = 1; j <= N; j++)
            if (j <= k) s += (r * (double)j - (double)(k - j)) * a.c[j] * p.c[j <= k ? k - j : 0];
        p.c[k] = s / ((double)k * a.c[0]);
    }
    return p;
}
template <int N>
PN_HD PnTaylor<N> pn_pow_half(const PnTaylor<N>& a, int k) { return pn_pow(a, 0.5 * (double)k); }
template <int N>
PN_HD PnTaylor<N> pn_exp(const PnTaylor<N>& a) {
    PnTaylor<N> e;
    e.c[0] = exp(a.c[0]);
#pragma unroll
    for (int k = 1; k <= N; k++) {
        double s = 0.0;
#pragma unroll
        for (int j = 1; j <= N; j++)
            if (j <= k) s += (double)j * a.c[j] * e.c[j <= k ? k - j : 0];
        e.c[k] = s / (double)k;
    }
    return e;
}
template <int N>
PN_HD PnTaylor<N> pn_log(const PnTaylor<N>& a) {
    PnTaylor<N> l;
    l.c[0] = log(a.c[0]);
#pragma unroll
    for (int k = 1; k <= N; k++) {
        double s = 0.0;
#pragma unroll
        for (int j = 1; j <= N; j++)
            if (j < k) s += (double)j * l.c[j] * a.c[j < k ? k - j : 0];
        l.c[k] = (a.c[k] - s / (double)k) / a.c[0];
    }
    return l;
}
template <int N>
PN_HD void pn_sincos(const PnTaylor<N>& a, PnTaylor<N>& s, PnTaylor<N>& c) {
    s.c[0] = sin(a.c[0]);
    c.c[0] = cos(a.c[0]);
#pragma unroll
    for (int k = 1; k <= N; k++) {
        double ss = 0.0, cc = 0.0;
#pragma unroll
        for (int j = 1; j <= N; j++) {
            if (j > k) continue;
            ss += (double)j * a.c[j] * c.c[k - j];
            cc -= (double)j * a.c[j] * s.c[k - j];
        }
        s.c[k] = ss / (double)k;
        c.c[k] = cc / (double)k;
    }
}
template <int N>
PN_HD PnTaylor<N> pn_sin(const PnTaylor<N>& a) {
    PnTaylor<N> s, c;
    pn_sincos(a, s, c);
    return s;
}
template <int N>
PN_HD PnTaylor<N> pn_cos(const PnTaylor<N>& a) {
    PnTaylor<N> s, c;
    pn_sincos(a, s, c);
    return c;
}
template <int N>
PN_HD PnTaylor<N> pn_tanh(const PnTaylor<N>& a) {
    // t' = (1 - t^2) a'
    PnTaylor<N> t, w;
    t.c[0] = tanh(a.c[0]);
    w.c[0] = 1.0 - t.c[0] * t.c[0];
#pragma unroll
    for (int k = 1; k <= N; k++) {
        double s = 0.0;
#pragma unroll
        for (int j = 1; j <= N; j++)
            if (j <= k) s += (double)j * a.c[j] * w.c[j <= k ? k - j : 0];
        t.c[k] = s / (double)k;
        double ww = 0.0;
#pragma unroll
        for (int j = 0; j <= N; j++)
            if (j <= k) ww -= t.c[j] * t.c[j <= k ? k - j : 0];
        w.c[k] = ww;
    }
    return t;
}

// Exact initial derivatives out[k*d + i] = d^k u_i/dt^k (t0), k = 0..Q, by the jet recursion
// u_{k+1} = f(u)_k / (k + 1) (same fixed point as TaylorIntegration.jetcoeffs!).
template <class Prob, int Q>
PN_HD void pn_taylor_init(double* out, const double* u0, const double* p, double t0) {
    constexpr int d = Prob::d;
    PnTaylor<Q> U[d], dU[d], tT;
    tT.c[0] = t0;
    if (Q >= 1) tT.c[Q >= 1 ? 1 : 0] = 1.0;
#pragma unroll
    for (int i = 0; i < d; i++) U[i] = PnTaylor<Q>(u0[i]);
    for (int ord = 0; ord < Q; ord++) {
        Prob::template f<PnTaylor<Q>>(dU, U, p, tT);
#pragma unroll
        for (int i = 0; i < d; i++) {
#pragma unroll
            for (int k = 0; k < Q; k++)
                if (k == ord) U[i].c[k + 1] = dU[i].c[k] / (double)(k + 1);
        }
    }
    double fact = 1.0;
#pragma unroll
    for (int k = 0; k <= Q; k++) {
        if (k > 0) fact *= (double)k;
#pragma unroll
        for (int i = 0; i < d; i++) out[k * d + i] = U[i].c[k] * fact;
    }
}

#endif /* PN_TAYLOR_CUH */
